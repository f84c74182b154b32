#!/bin/bash
# GAE kernel configurations for mid-size batches (which tile / staged-output variant per B)
for B in 4096 8192 16384 32768; do
  B=$B CFGS=0,1,2,4,5,13,21,31,35,36,37,38,30,34 python tools/tune_gae.py 2>&1 | grep -v copy_ms | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('B=%6d cfg %2d fwd %.4f ms (%.0f GB/s) bwd %.4f ms (%.0f GB/s)'%(r['B'],r['cfg'],r['fwd_ms'],r['fwd_gbs'],r['bwd_ms'],r['bwd_gbs']))
"
done
