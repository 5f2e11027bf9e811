import json, sys
for f in sys.argv[1:]:
    try:
        d = None
        for l in open(f):
            l = l.strip()
            if l.startswith('{'):
                d = json.loads(l)
        if d is None:
            print(f, "NO JSON"); continue
        r = d.get("roofline") or {}
        st = r.get("step", {})
        km = st.get("kernel_ms", {})
        print(f"{f.split('/')[-1]:28s} ms/step {d['ms_per_step']:.4f} value {d['value']:.0f} e2e {d['e2e']['value']:.0f} "
              f"mac_ms {r.get('mean_launch_ms',0):.4f} frac {r.get('frac',0):.3f} stepfrac {st.get('frac',0):.3f} "
              f"k1 {km.get('k_r2c_ingest',0):.4f} k2 {km.get('k_c2r_emit',0):.4f} plan {r.get('launch_plan')} clk {d.get('clocks',{}).get('sm_mhz')}")
    except Exception as e:
        print(f, "ERR", e)
