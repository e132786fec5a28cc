#!/bin/bash
# tuning sweep: the device-resident bench line with alternative builds of the library (KMAP_B200_LIB)
mkdir -p gpurun_out
for lib in kmap_b200/libkmap_b200.so gpurun_variants_*.so; do
  KMAP_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-e2e --no-cpu --no-hamdist --no-piece2 --no-workflow --no-crosscheck --steps 5 --warmup 3 > gpurun_out/sweep_tmp.log 2>/dev/null
  python - "$lib" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/sweep_tmp.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); p=d['roofline']['phases_ms']
    print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in p.items() if k in ('dedup_scan','bucket_hist','partition','bucket_count')}, d['checks'])
else:
    print(sys.argv[1], "FAILED")
PY
done
