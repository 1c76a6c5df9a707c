mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "k1_dense or end_to_end or k3_on_demand" 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --cpu-pairs 0 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'matched', d['matched_fraction'], 'rmse', d['pose_rmse_px_vs_ground_truth']); print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if k.startswith('conv')})"
