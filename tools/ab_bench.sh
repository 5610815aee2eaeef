#!/bin/bash
# A/B of an environment switch on the same box: tools/ab_bench.sh VAR  (runs bench twice each, alternating)
V=$1
for i in 1 2; do
  for mode in off on; do
    if [ $mode = on ]; then export $V=1; else unset $V; fi
    timeout 300 python bench.py --no-cpu-baseline --engines ${ENGINES:-1} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V=$mode', 'value %.1f e2e %.1f frac %.4f conv_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_ms_per_step']['conv']), d['clocks'])"
  done
done
