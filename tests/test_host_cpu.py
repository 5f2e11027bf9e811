"""
CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/pgx.h
declares (no compute calls), the host-side mirror of the PE protocol behaves like the reference
(restated reference tests: tests/test_extent.py, test_snippet.py, test_mix_pe.py, test_spatial_pe.py,
test_renderer.py, test_convolve_pe.py:30-42), the product path fails loudly without a CUDA device,
the numpy model of the device schedule agrees with the oracle, and world_size-2 gloo sharding works.
"""
import os
import re

import numpy as np
import pytest

import pygmu2_b200 as pg
from pygmu2_b200 import _lib, dist as pdist, kemar
from conftest import ROOT, golden


# ---- C ABI -------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgx.h")).read()
    declared = set(re.findall(r"PGX_API\s+(?:const\s+char\*|void\*|int)\s+(pgx_\w+)\s*\(", hdr))
    assert declared, "no prototypes parsed from pgx.h"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    h = _lib.lib()
    for name in declared:
        assert hasattr(h, name), name
    assert h.pgx_abi_version() == _lib.ABI_VERSION
    assert isinstance(h.pgx_last_error(), bytes)
    # the entry-point count quoted in the documents is the header's
    for doc in ("DESIGN.md", "INTEGRATION.md"):
        txt = open(os.path.join(ROOT, doc)).read()
        assert f"{len(declared)} entry points" in txt or f"all {len(declared)} prototypes" in txt, doc


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.Layout) == 24
    assert C.sizeof(_lib.BankConfig) == 44
    assert C.sizeof(_lib.Profile) == 72
    assert C.sizeof(_lib.OscConfig) == 40


def test_integration_md_stub_matches_the_library():
    """The reference-side ctypes stub printed in INTEGRATION.md is executed as written against the built library:
    its structs must have the sizes the library was compiled with (round 1 shipped a stub without tail_block)."""
    import ctypes as C
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    stub = next(b for b in blocks if "class BankConfig" in b)
    os.environ["PYGMU2_PGX"] = _lib.LIB_PATH
    ns = {}
    try:
        exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    finally:
        del os.environ["PYGMU2_PGX"]
    h = _lib.lib()
    assert C.sizeof(ns["Layout"]) == h.pgx_struct_size(0) == C.sizeof(_lib.Layout)
    assert C.sizeof(ns["BankConfig"]) == h.pgx_struct_size(1) == C.sizeof(_lib.BankConfig)
    assert [f[0] for f in ns["BankConfig"]._fields_] == [f[0] for f in _lib.BankConfig._fields_]
    hdr = open(os.path.join(ROOT, "include", "pgx.h")).read()
    body = re.search(r"typedef struct pgx_bank_config \{(.*?)\} pgx_bank_config;", hdr, flags=re.S).group(1)
    fields = re.findall(r"^\s*u?int32_t\s+(\w+);", body, flags=re.M)
    assert fields == [f[0] for f in ns["BankConfig"]._fields_]
    assert h.pgx_struct_size(99) == -1


def test_no_silent_cpu_path():
    """Without a CUDA device the device-backed PEs raise; nothing falls back to numpy."""
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    pe = pg.ConvolvePE(pg.ArrayPE([1.0, 2.0, 3.0]), pg.ArrayPE([1.0, 1.0]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pe.render(0, 3)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pg.device_mix_sum([np.zeros((4, 1), np.float32)] * 2)
    sp = pg.SpatialPE(pg.ArrayPE(np.ones(16)), method=pg.SpatialHRTF(30.0))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        sp.render(0, 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pygmu2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pygmu2_oracle" not in src and "import oracle" not in src, f


# ---- Extent / Snippet (reference tests/test_extent.py, tests/test_snippet.py) ---------------------
def test_extent_algebra():
    E = pg.Extent
    with pytest.raises(ValueError):
        E(5, 2)
    assert E(0, 10).contains(0) and not E(0, 10).contains(10)
    assert E(None, None).contains(-10 ** 9)
    assert E(0, 10).spans(2, 8) and not E(0, 10).spans(2, 9)
    assert E(0, 10).intersects(E(9, 20)) and not E(0, 10).intersects(E(10, 20))
    assert not E(5, 5).intersects(E(0, 10)) and not E(5, 5).intersects(E(5, 5))
    assert E(0, 10).intersection(E(5, 20)) == E(5, 10)
    assert E(0, 10).intersection(E(20, 30)).is_empty()
    assert E(0, 100).union(E(50, 200)) == E(0, 200)
    assert E(0, 100).union(E(None, None)) == E(None, None)
    assert E(0, 10).union(E(3, 3)) == E(0, 10)
    assert E(None, 5).duration is None and E(2, 5).duration == 3
    assert not E(3, 3) and E(0, 1)


def test_snippet_rules():
    s = pg.Snippet(7, np.arange(4, dtype=np.float64))
    assert s.data.shape == (4, 1) and s.data.dtype == np.float32 and (s.start, s.end, s.duration) == (7, 11, 4)
    with pytest.raises(ValueError):
        pg.Snippet(0, np.zeros((2, 2, 2)))
    z = pg.Snippet.from_zeros(3, 0, 2)
    assert z.duration == 0 and z.channels == 2
    assert pg.Snippet(0, np.ones((3, 2))) == pg.Snippet(0, np.ones((3, 2)) + 1e-9)


def test_sample_rate_required_before_construction():
    pg.core._state["sample_rate"] = None
    with pytest.raises(RuntimeError):
        pg.ArrayPE([1.0])
    pg.set_sample_rate(44100)


def test_render_wrapper_contract():
    a = pg.ArrayPE([1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        a.render(0, -1)
    assert a.render(5, 0).data.shape == (0, 1)
    np.testing.assert_array_equal(a.render(-1, 5).data[:, 0], [0, 1, 2, 3, 0])  # zero fill outside extent


# ---- ConvolvePE host logic (reference tests/test_convolve_pe.py:30-42,165-185) ---------------------
def test_convolve_filter_extent_contract():
    src = pg.ArrayPE([1, 2, 3, 4])
    with pytest.raises(ValueError):
        pg.ConvolvePE(src, pg.CropPE(pg.ArrayPE([1, 0, 0]), 1, 2)).extent()   # start != 0
    with pytest.raises(ValueError):
        pg.ConvolvePE(src, pg.ConstantPE(1.0)).extent()                        # infinite


def test_convolve_extent_channels_props():
    pe = pg.ConvolvePE(pg.ArrayPE(np.zeros((10, 1))), pg.ArrayPE(np.zeros((4, 2))), fft_size=64)
    assert pe.extent() == pg.Extent(0, 13)
    assert pe.channel_count() == 2 and pe.fft_size == 64 and not pe.is_pure()
    assert pe.inputs() == [pe.src, pe.fir]
    assert pg.ConvolvePE(pg.ArrayPE(np.zeros((10, 2))), pg.ArrayPE(np.zeros(4))).channel_count() == 2
    assert pg.ConvolvePE(pg.SinePE(), pg.ArrayPE(np.zeros(4))).extent() == pg.Extent(None, None)
    assert "ConvolvePE(src=ArrayPE, fir=ArrayPE, fft_size=64)" == repr(pe)


def test_ir_energy_norm():
    assert pg.ConvolvePE.ir_energy_norm(pg.ArrayPE([3.0, 4.0])) == pytest.approx(5.0, rel=1e-5)
    assert pg.ConvolvePE.ir_energy_norm(pg.ArrayPE([1.0])) == pytest.approx(1.0, rel=1e-5)
    assert pg.ConvolvePE.ir_energy_norm(pg.ArrayPE([0.0, 0.0])) == 1.0
    assert pg.ConvolvePE.ir_energy_norm(pg.ArrayPE(np.ones((2, 2)))) == pytest.approx(2.0, rel=1e-5)
    assert pg.ConvolvePE.ir_energy_norm(pg.ConstantPE(1.0)) == 1.0


# ---- MixPE host logic (reference tests/test_mix_pe.py:17-64,161-227) --------------------------------
def test_mix_construction_extent_channels():
    with pytest.raises(ValueError, match="at least 2 inputs"):
        pg.MixPE(pg.ConstantPE(1.0))
    m = pg.MixPE([pg.ConstantPE(0.3), pg.ConstantPE(0.4)])
    assert len(m.inputs()) == 2 and m.is_pure() and repr(m) == "MixPE(ConstantPE, ConstantPE)"
    assert m.extent() == pg.Extent(None, None)
    assert pg.MixPE(pg.ArrayPE(np.zeros(100)), pg.DelayPE(pg.ArrayPE(np.zeros(150)), 50)).extent() == pg.Extent(0, 200)
    assert pg.MixPE(pg.ArrayPE(np.zeros(100)), pg.ConstantPE(0.0)).extent() == pg.Extent(None, None)
    assert pg.MixPE(pg.ConstantPE(0, 2), pg.ConstantPE(0, 2)).channel_count() == 2
    with pytest.raises(ValueError, match="channel mismatch"):
        m.resolve_channel_count([1, 2])
    out = pg.MixPE(pg.ArrayPE(np.ones(4)), pg.ArrayPE(np.ones(4))).render(100, 8)   # every input misses
    assert out.data.shape == (8, 1) and not out.data.any()


# ---- Spatial host logic (reference tests/test_spatial_pe.py) ------------------------------------------
def test_spatial_methods_host():
    mono = pg.Snippet(0, np.arange(1, 5, dtype=np.float32))
    st = pg.Snippet(0, np.array([[1, 3], [2, 4]], dtype=np.float32))
    assert np.array_equal(pg.SpatialAdapter(2).render(mono, 0, 4, 44100), np.repeat(mono.data, 2, axis=1))
    assert np.array_equal(pg.SpatialAdapter(1).render(st, 0, 2, 44100)[:, 0], [2, 3])
    q = pg.SpatialAdapter(4).render(st, 0, 2, 44100)
    assert np.array_equal(q, [[1, 3, 2, 2], [2, 4, 3, 3]])
    with pytest.raises(ValueError):
        pg.SpatialAdapter(0)
    lin = pg.SpatialLinear(azimuth=0.0).render(mono, 0, 4, 44100)
    np.testing.assert_allclose(lin[:, 0], 0.5 * mono.data[:, 0], atol=1e-6)
    cp = pg.SpatialConstantPower(azimuth=200.0).render(mono, 0, 4, 44100)     # clamped to +90: hard right
    np.testing.assert_allclose(cp[:, 0], 0.0, atol=1e-6)
    np.testing.assert_allclose(cp[:, 1], mono.data[:, 0], atol=1e-6)
    c0 = pg.SpatialConstantPower(azimuth=0.0).render(mono, 0, 4, 44100)
    np.testing.assert_allclose(c0[:, 0] ** 2 + c0[:, 1] ** 2, mono.data[:, 0] ** 2, rtol=1e-5)


def test_spatial_pe_contract():
    with pytest.raises(ValueError):
        pg.SpatialPE(pg.SinePE(), method=None)
    az = pg.ConstantPE(10.0)
    sp = pg.SpatialPE(pg.ArrayPE(np.zeros(10)), method=pg.SpatialLinear(azimuth=az))
    assert sp.inputs()[1] is az and sp.extent() == pg.Extent(0, 10) and sp.is_pure() and sp.channel_count() == 2
    with pytest.raises(ValueError, match="must be static"):
        pg.SpatialHRTF(azimuth=az)
    with pytest.raises(ValueError, match="must be static"):
        pg.SpatialHRTF(azimuth=0.0, elevation=az)
    m = pg.SpatialHRTF(azimuth=45, elevation=10)
    assert m.azimuth == 45.0 and m.output_channels == 2 and repr(m) == "SpatialHRTF(azimuth=45.0, elevation=10.0)"


def test_hrtf_filename_lookup_and_table():
    f = pg.SpatialHRTF.hrtf_filename_for
    assert f(0, 0) == "H0e000a.wav" and f(45, 0) == "H0e045a.wav" and f(-45, 0) == f(45, 0)
    assert f(90, 0) == "H0e090a.wav" and f(0, 30).startswith("H30")
    names = {e[2] for e in pg.SpatialHRTF.KEMAR_HRTF_ENTRIES}
    assert len(pg.SpatialHRTF.KEMAR_HRTF_ENTRIES) == 368 and f(123.4, 56.7) in names
    g = golden("hrtf_lookup.npz")    # picks of the real reference, incl. ties and clamping
    idx = np.array([kemar.nearest_index(a, e) for a, e in zip(g["az"], g["el"])])
    assert np.array_equal(idx, g["idx"])
    table, sr = kemar.load_table()
    assert table.shape == (368, 128, 2) and table.dtype == np.float32 and sr == 44100


# ---- Renderer (reference tests/test_renderer.py) -----------------------------------------------------
class _Stateful(pg.ProcessingElement):
    def __init__(self, src):
        self._src, self.log = src, []

    def inputs(self):
        return [self._src]

    def _on_start(self):
        self.log.append("start")

    def _on_stop(self):
        self.log.append("stop")

    def _render(self, start, duration):
        return self._src.render(start, duration)


def test_renderer_contract():
    r = pg.NullRenderer(sample_rate=44100)
    with pytest.raises(RuntimeError):
        r.start()
    pe = _Stateful(pg.ConstantPE(0.5))
    r.set_source(pe)
    with pytest.raises(RuntimeError):
        r.render(0, 10)
    r.start()
    with pytest.raises(ValueError):
        r.render(0, 0)
    r.render(0, 16)
    with pytest.raises(RuntimeError):
        r.set_source(pe)
    r.stop()
    r.stop()
    assert pe.log == ["start", "stop"] and r.channel_count == 1
    shared = _Stateful(pg.ConstantPE(0.1))
    with pytest.raises(ValueError, match="not pure"):
        pg.NullRenderer().set_source(pg.MixPE(shared, shared, fuse=False))
    with pg.NullRenderer() as r2:
        r2.set_source(pe)
        r2.start()
    assert not r2.started


# ---- device schedule model vs oracle ---------------------------------------------------------------
def test_kernel_model_matches_oracle():
    import kernel_model as km
    import pygmu2_oracle as orc
    rng = np.random.default_rng(3)
    for L, B, vec in ((3, 16, False), (70, 16, False), (900, 64, True)):
        h = (rng.standard_normal(L) / np.sqrt(L)).astype(np.float32)
        x = rng.uniform(-1, 1, 260 if not vec else 2500).astype(np.float32)
        mb, o = km.ModelBank(h, B, vec), orc.OracleConvolve(h, 1)
        ys, yr, pos = [], [], 0
        for d in (1, 17, B - 1, B, B + 1, 3 * B + 5, 10 ** 6):
            d = min(d, len(x) - pos)
            if d <= 0:
                break
            ys.append(mb.process(x[pos:pos + d]))
            yr.append(o.render(x[pos:pos + d])[:, 0])
            pos += d
        ys, yr = np.concatenate(ys), np.concatenate(yr)
        assert np.max(np.abs(ys - yr)) <= 1e-5 * np.max(np.abs(yr))


def test_choose_block_and_workload_bytes():
    from pygmu2_b200 import workloads as wl
    assert pg.choose_block(132300, 512) == 512 and pg.choose_block(4096, 441000) == 4096
    assert pg.choose_block(3, 6) == 16 and pg.choose_block(100, 64) == 128
    # B is sticky from the first pull: a probing render(0, 1) must not lock a long filter into thousands of tiny
    # partitions (floor of 256, at most 2048 partitions); block_size= pins anything smaller (the named C5 does)
    assert pg.choose_block(132300, 1) == 256 and pg.choose_block(441000, 64) == 256
    assert pg.choose_block(3_000_000, 16) == 2048
    # bytes per output sample.channel
    assert round(wl.bytes_per_block_step(256, 2, 2, 132300, 512, False) / (256 * 2 * 512)) == 2092
    assert round(wl.bytes_per_block_step(256, 2, 2, 132300, 512, True) / (256 * 2 * 512)) == 4168
    assert round(wl.bytes_per_block_step(512, 1, 1, 88200, 512, True) / (512 * 512)) == 2789


# ---- multi-process sharding over gloo (world_size 2) -----------------------------------------------
def test_shard_bounds_partition():
    for n, w in ((4096, 8), (10, 3), (2, 4), (7, 7)):
        spans = [pdist.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _gloo_worker(rank, world, port, ret):
    import sys
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import pygmu2_oracle as orc
    from pygmu2_b200 import dist as pd, workloads as wl
    pd.init_process_group("gloo")
    N, L, n = 6, 700, 1024
    sm = None

    def local_mix(x_local):  # oracle stands in for the device bank: this test covers the host plumbing
        parts = [orc.OracleConvolve(wl.c4_ir(sm.lo + i, L), 1).render(x_local[i, 0])[:, 0] for i in range(x_local.shape[0])]
        return np.sum(np.stack(parts), axis=0, dtype=np.float32)[None, :]

    sm = pd.ShardedMix(N, local_mix=local_mix, root=0)
    x_local = np.stack([wl.c4_input(n, s) for s in range(sm.lo, sm.hi)])[:, None, :]
    y = sm.render_mix_host(x_local)
    if rank == 0:
        ret["y"] = y.copy()
        ret["span0"] = (sm.lo, sm.hi)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mix_gloo_world2():
    import torch.multiprocessing as mp
    import pygmu2_oracle as orc
    from pygmu2_b200 import workloads as wl
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    N, L, n = 6, 700, 1024
    full = np.sum(np.stack([orc.OracleConvolve(wl.c4_ir(s, L), 1).render(wl.c4_input(n, s))[:, 0] for s in range(N)]),
                  axis=0, dtype=np.float64)
    assert ret["span0"] == (0, 3)
    assert np.max(np.abs(ret["y"][0] - full)) <= 1e-5 * np.max(np.abs(full))


# ---- WAV staging: host formulas and file round trip ------------------------
def test_pcm16_conversion_rules_and_wav_round_trip(tmp_path):
    from pygmu2_b200.wav_pe import f32_to_pcm16, pcm16_to_f32
    x = np.array([0.0, 1.0, -1.0, 2.0, -2.0, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, -0.5 / 32768, -1.5 / 32768,
                  32766.6 / 32768, 32767.0 / 32768, -32767.5 / 32768], dtype=np.float32)
    want = np.array([0, 32767, -32768, 32767, -32768, 0, 2, 2, 0, -2, 32767, 32767, -32768], dtype=np.int16)
    np.testing.assert_array_equal(f32_to_pcm16(x), want)          # clip + round-half-even (libsndfile f2s_clip_array)
    p = np.arange(-32768, 32768, 7, dtype=np.int16)
    np.testing.assert_array_equal(f32_to_pcm16(pcm16_to_f32(p)), p)   # int16 -> float32 -> int16 is the identity
    pg.set_sample_rate(22_050)
    rng = np.random.default_rng(2)
    data = rng.uniform(-1.2, 1.2, (1000, 2)).astype(np.float32)       # includes samples that clip
    path = str(tmp_path / "rt.wav")
    w = pg.WavWriterPE(pg.ArrayPE(data), path)
    r = pg.NullRenderer(sample_rate=22_050)
    r.set_source(w)
    with r:
        r.start()
        r.render(0, 600)
        r.render(600, 400)
    assert w.frames_written == 1000
    rd = pg.WavReaderPE(path)
    assert rd.extent() == pg.Extent(0, 1000) and rd.channel_count() == 2 and rd.sample_rate == 22_050
    got = rd.render(-10, 1020).data
    assert not got[:10].any() and not got[1010:].any()
    np.testing.assert_array_equal(got[10:1010], pcm16_to_f32(f32_to_pcm16(data)))
    np.testing.assert_array_equal(rd.render_pcm16(0, 1000), f32_to_pcm16(data))


# ---- device-source PEs: host-side parameter logic against the oracle restatement (no GPU needed) -------
def test_supersaw_host_parameters_match_reference_rules():
    import pygmu2_oracle_sources as osrc
    g = golden("src_oscillators.npz")
    mixname = {0: "center_heavy", 1: "linear", 2: "equal"}
    for f, a, v, d, mm, rp, sd in g["ssaw_cases"]:
        pe = pg.SuperSawPE(frequency=f, amplitude=a, voices=int(v), detune_cents=d, mix_mode=mixname[int(mm)],
                           randomize_phase=bool(rp), seed=int(sd))
        fr, ga, ph = osrc.supersaw_params(f, int(v), d, mixname[int(mm)], bool(rp), int(sd))
        np.testing.assert_array_equal(np.array(pe._osc_freq), fr)
        np.testing.assert_array_equal(np.array(pe._osc_gain), ga)
        np.testing.assert_array_equal(np.array(pe._osc_phase), ph)
        assert pe.channel_count() == 1 and not pe.is_pure() and pe.extent() == pg.Extent(None, None)
    with pytest.raises(ValueError):
        pg.SuperSawPE(440.0, mix_mode="nope")
    assert pg.SinePE().is_pure() and pg.SinePE(channels=2).channel_count() == 2
    assert pg.BlitSawPE(100.0, initial_phase=1.25).initial_phase == 0.25 and pg.BlitSawPE(100.0).m is None


def test_mix_unwrap_folds_delay_and_gain_wrappers():
    from pygmu2_b200.mix_pe import _unwrap, _foldable_method
    core = pg.SpatialPE(pg.ArrayPE(np.zeros(8, np.float32)), method=pg.SpatialHRTF(10.0))
    pe = pg.DelayPE(pg.GainPE(pg.DelayPE(pg.GainPE(core, 0.5), 100), 0.25), 7)
    c, d, gn = _unwrap(pe)
    assert c is core and d == 107 and gn == np.float32(0.125)
    assert _unwrap(core) == (core, 0, None)
    dyn = pg.GainPE(core, gain=pg.ConstantPE(0.5))                    # a PE-valued gain is not folded
    assert _unwrap(dyn)[0] is dyn
    assert _foldable_method(pg.SpatialLinear(30.0)) and _foldable_method(pg.SpatialConstantPower(-30.0))
    assert not _foldable_method(pg.SpatialLinear(pg.ConstantPE(0.0))) and not _foldable_method(pg.SpatialAdapter(2))


def test_vectorised_hrtf_lookup_equals_the_reference_rule():
    """kemar.nearest_indices (4 candidates per elevation ring) must pick exactly the entry of the full
    first-minimum scan (spatial_pe.py:395-426), ties, clamping and out-of-range elevations included."""
    g = golden("hrtf_lookup.npz")
    np.testing.assert_array_equal(kemar.nearest_indices(g["az"], g["el"]), g["idx"])
    rng = np.random.default_rng(9)
    az = np.concatenate([rng.uniform(-250, 250, 5000), np.arange(-180, 181, 0.5), np.repeat(np.arange(0, 181, 2.5), 2)])
    el = np.concatenate([rng.uniform(-90, 130, 5000), np.tile([0.0, 5.0, -45.0, 85.0, 95.0], 145)[:721],
                         np.tile([15.0, 45.0], 73)])
    n = min(az.shape[0], el.shape[0])
    az, el = az[:n], el[:n]
    full = np.array([kemar.nearest_index(a, e) for a, e in zip(az, el)])
    np.testing.assert_array_equal(kemar.nearest_indices(az, el), full)


def test_fft16_register_model_matches_numpy_fft():
    """The index maps of the radix-16 kernel (thread digit <-> register exchanges, digit-reversed butterfly outputs,
    twiddle exponents), as modelled in tests/kernel_model.py, are an exact 4096-point DFT / inverse DFT."""
    import kernel_model as km
    rng = np.random.default_rng(0)
    z = rng.standard_normal(km.FFT16_N) + 1j * rng.standard_normal(km.FFT16_N)
    assert np.max(np.abs(km.fft16_forward(z) - np.fft.fft(z))) < 1e-9
    assert np.max(np.abs(km.fft16_inverse(z) - np.fft.ifft(z) * km.FFT16_N)) < 1e-9


def test_nearest_direction_in_c_equals_numpy_and_the_scalar_rule():
    """pgx_nearest_direction (host C: the per-pull filter selection of moving sources) against the vectorised numpy
    search and the scalar restatement of spatial_pe.py:395-426, ties (grid points, midpoints, |az| > 180) included."""
    rng = np.random.default_rng(3)
    az = np.concatenate([rng.uniform(-220, 220, 4000), np.arange(-180, 181, 2.5), [0.0, 180.0, -180.0, 7.5, 172.5]])
    el = np.concatenate([rng.uniform(-60, 100, 4000), np.resize(np.arange(-40, 91, 5.0), 145), [0.0, 90.0, 85.0, -5.0, 45.0]])
    a = kemar.nearest_indices(az, el)
    assert np.array_equal(a, kemar.nearest_indices_numpy(az, el))
    assert np.array_equal(a[::37], [kemar.nearest_index(x, y) for x, y in zip(az[::37], el[::37])])


@pytest.mark.parametrize("P,T", [(2, 2), (5, 2), (5, 4), (8, 4), (33, 4), (33, 2)])
def test_time_tiled_pass_index_model_equals_the_partitioned_sum(P, T):
    """The index algebra of the time-tiled pass (tests/kernel_model.py restates issue_tile, k_fdl_mac_tile's two runs and
    sliding filter window, and k_c2r's recent rows): for every ring position, with passes issued every T blocks and
    re-issued on demand at arbitrary blocks, every block's output equals sum_p X[t-p] * H[p]."""
    import kernel_model as km
    rng = np.random.default_rng(P * 10 + T)
    K = 3
    R, n_spare, *_ = km.tiled_pass_terms(P, T, 0)
    H = rng.standard_normal((P, K)) + 1j * rng.standard_normal((P, K))
    Hd = np.zeros((2 * R, K), complex)
    for p in range(P):
        Hd[R - 1 - p] = Hd[2 * R - 1 - p] = H[p]
    n_blocks = 3 * R + 7
    X = rng.standard_normal((n_blocks, K)) + 1j * rng.standard_normal((n_blocks, K))
    ring = np.zeros((R, K), complex)
    base, S = None, None
    invalidate_at = set(rng.integers(1, n_blocks, 6).tolist())      # resets of the coverage (map change, reload, ...)
    for t in range(n_blocks):
        head = t % R
        if base is None or not (base <= t < base + T) or t in invalidate_at:
            S = km.tiled_pass(ring, Hd, P, T, head, n_split=1 + (t % 3))   # rows committed so far: blocks < t
            base = t
        ring[head] = X[t]                                            # K1 of block t
        y = km.tiled_output(ring, Hd, S[t - base], R, head, t - base)
        ref = sum(X[t - p] * H[p] for p in range(P) if t - p >= 0)
        np.testing.assert_allclose(y, ref, atol=1e-9, err_msg=f"block {t}")


def test_design_tables_match_the_committed_bench_records():
    """The measured tables of DESIGN.md are generated from profiles/*.json by scripts/design_tables.py: regenerating them
    must reproduce the committed text (a table edited by hand, or a refreshed record without its table, fails here)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("design_tables", os.path.join(root, "scripts", "design_tables.py"))
    dt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dt)
    text = open(os.path.join(root, "DESIGN.md")).read()
    for key, fn in (("BENCH_TABLE", dt.bench_table), ("MULTIGPU_NUMBERS", dt.multi_table), ("NAMED_TABLE", dt.named_table)):
        block = f"<!-- BEGIN {key} -->\n{fn()}\n<!-- END {key} -->"
        assert block in text, key


def test_committed_headline_record_carries_the_bench_contract():
    """The headline bench line as committed (profiles/r02c_bench_c2_k20.json: the driver-shaped default run on a B200) has
    every key of the bench contract, on the BASELINE metric and the named workload, with a green parity leg."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = [l for l in open(os.path.join(root, "profiles", "r02c_bench_c2_k20.json")).read().splitlines() if l.startswith("{")]
    d = json.loads(txt[-1])
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity"):
        assert k in d, k
    assert d["metric"].startswith("ConvolvePE audio-sec") and base["metric"].startswith("ConvolvePE audio-sec")
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert d["parity"]["ok"] and d["parity"]["max_rel_err"] <= 1e-5


def test_delay_pe_fractional_and_pe_valued_delays_match_the_reference_golden():
    """DelayPE's interpolated modes (delay_pe.py:163-228, interpolated_lookup.py:89-145) against outputs of the REAL
    reference (oracle/gen_golden_delay.py -> tests/golden/delay_interp.npz): fractional delays of either sign and a
    PE-valued vibrato, linear and cubic, ragged pulls from before the source starts to after it ends, and the extents."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_golden_delay", os.path.join(root, "oracle", "gen_golden_delay.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(root, "tests", "golden", "delay_interp.npz"))
    pg.set_sample_rate(44_100)
    cases = gen.cases(pg, lambda n: pg.InterpolationMode(n))
    assert len(cases) == 6
    for name, pe in cases.items():
        y = gen.pull(pe)
        assert y.dtype == np.float32 and y.shape == gold[name].shape
        np.testing.assert_array_equal(y, gold[name], err_msg=name)          # same arithmetic in numpy: bit for bit
        e = pe.extent()
        assert [e.start, e.end] == gold[name + "_extent"].tolist(), name
    # whole-number floats are integer delays (delay_pe.py:72-78) and stay foldable into a fused mix
    assert pg.DelayPE(pg.ArrayPE(np.ones(8)), 4.0).mode == "int" and pg.DelayPE(pg.ArrayPE(np.ones(8)), 4.5).mode == "float"
    ctl = pg.DelayPE(pg.ArrayPE(np.ones(8)), pg.ConstantPE(2.0))
    assert ctl.mode == "pe" and len(ctl.inputs()) == 2
