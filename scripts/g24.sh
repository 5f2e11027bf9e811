cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in prev new; do
  if [ $v = prev ]; then export PGX_LIB=$GRAFT_REPO_ROOT/ab/libpgx_prev.so; else unset PGX_LIB; fi
  for w in c2 c1; do
    timeout 300 python bench.py --steps 1000 --warmup 20 --workload $w --no-cpu > gpurun_out/g24_${w}_${v}_$rep.json 2> gpurun_out/g24_${w}_${v}_$rep.err
  done
done
done
