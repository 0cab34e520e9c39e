mkdir -p gpurun_out
for pf in 0 16384 65536 262144; do
  echo "PF=$pf" >> gpurun_out/r09_pf.log
  MFLBM_PF_DIST=$pf timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus\|"kernel_ms_per_step": [0-9.]*' >> gpurun_out/r09_pf.log
  MFLBM_PF_DIST=$pf timeout 600 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus\|"kernel_ms_per_step": [0-9.]*' >> gpurun_out/r09_pf.log
done
cat gpurun_out/r09_pf.log
