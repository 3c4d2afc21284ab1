for cfg in "1:1" "1:" "1,2:" "2:" ":"; do
  sp=${cfg%%:*}; so=${cfg##*:}
  SPVD_SP_LEVELS=$sp SPVD_SORT_LEVELS=$so python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('SP=$sp SORT=$so', d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
done
