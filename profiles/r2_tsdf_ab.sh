set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tsdf_kernel_variants" 2>&1 | tail -15
for v in "NVB_TSDF_TMA=0" "NVB_TSDF_TMA=1 NVB_TSDF_TMA_CTAS_PER_SM=1" "NVB_TSDF_TMA=1 NVB_TSDF_TMA_CTAS_PER_SM=2" "NVB_TSDF_TMA=1 NVB_TSDF_TMA_CTAS_PER_SM=4" "NVB_TSDF_TMA=1 NVB_TSDF_TMA_CTAS_PER_SM=8"; do
  echo "== $v"
  env $v python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['value'], {k:round(v['ms_per_frame']*1000,1) for k,v in j['stages'].items()}, j.get('parity_checked'))
"
done
