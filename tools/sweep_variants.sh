for v in b512x2 b256x3 b384x2 b256x2 b128x5 b1024x1; do
  if [ $v = b512x2 ]; then lib=stormwater-management-model_b200/csrc/libswmm_b200.so; else lib=variants/libswmm_b200_$v.so; fi
  python tools/bench_variant.py $lib --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '%.3e'%d['value'], {k:round(v) for k,v in d['roofline']['phase_ms'].items()})"
done
