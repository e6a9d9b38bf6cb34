# usage: bash tools/sweep_vlib2.sh <members> name1 name2 ...   (vlib/libswmm_b200_<name>.so; "main" = the product library;
# a name with suffix @persist runs with SWB_STAGED_MIN_M=0)
M=$1; shift
for v in "$@"; do
  name=${v%@persist}
  lib=vlib/libswmm_b200_$name.so
  [ "$name" = main ] && lib=stormwater-management-model_b200/csrc/libswmm_b200.so
  env=""
  [ "$v" != "$name" ] && export SWB_STAGED_MIN_M=0 || unset SWB_STAGED_MIN_M
  python tools/bench_variant.py $lib --members $M --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak $EXTRA 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$EXTRA', 'M=$M', '%.3e'%d['value'], 'launches', d['gpu_launches'], {k:round(v) for k,v in d['roofline']['phase_ms'].items() if v})"
done
