#!/bin/bash
# quick GPU check: parity tests + the MSM/NTT numbers of the bench line
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('MSM pts/s %.4g  ms/step %.2f  acc_ms %.2f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value']))
print({k:round(v['ms'],4) for k,v in d.get('extra',{}).items()})"
