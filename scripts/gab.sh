cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
for v in prev cur; do
  if [ $v = cur ]; then unset PGX_LIB; else export PGX_LIB=$GRAFT_REPO_ROOT/ab/libpgx_$v.so; fi
  for w in c1 c2 c3; do
  timeout 300 python bench.py --steps 1500 --warmup 20 --workload $w --no-cpu > gpurun_out/gab_${w}_${v}_$rep.json 2> gpurun_out/gab_${w}_${v}_$rep.err
  done
done
done
