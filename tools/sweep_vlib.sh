# usage: bash tools/sweep_vlib.sh name1 name2 ...   (vlib/libswmm_b200_<name>.so; "main" = the product library)
for v in "$@"; do
  lib=vlib/libswmm_b200_$v.so
  [ "$v" = main ] && lib=stormwater-management-model_b200/csrc/libswmm_b200.so
  python tools/bench_variant.py $lib --members 512 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-c5 --no-c2-single --no-weak 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '%.3e'%d['value'], {k:round(v) for k,v in d['roofline']['phase_ms'].items()})"
done
