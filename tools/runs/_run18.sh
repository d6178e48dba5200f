timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "batchnorm or bn or host_pipeline" > gpurun_out/r2_t14.log 2>&1; tail -3 gpurun_out/r2_t14.log
pr() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['ms_per_step_p10_p50_p90'], d['value'], d['e2e']['value'], d['exposed_push_pull_ms_per_step'], d['gpu_launches_per_step'], d['protocol_errors'])"; }
GEOMX_STEP_OVERLAP=0 GEOMX_STEP_EXCHANGE=sharded python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b18_n2_sh.log 2>&1; pr gpurun_out/r2_b18_n2_sh.log N2-single-sharded
GEOMX_STEP_OVERLAP=0 GEOMX_STEP_EXCHANGE=replicated python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b18_n2_rep.log 2>&1; pr gpurun_out/r2_b18_n2_rep.log N2-single-replicated
GEOMX_STEP_OVERLAP=0 python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b18_n2_ll.log 2>&1; pr gpurun_out/r2_b18_n2_ll.log N2-single-ll
GEOMX_DENSE_CHANNEL_GRID=16 python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b18_n2_g16.log 2>&1; pr gpurun_out/r2_b18_n2_g16.log N2-channels-grid16
GEOMX_DENSE_CHANNEL_GRID=40 python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b18_n2_g40.log 2>&1; pr gpurun_out/r2_b18_n2_g40.log N2-channels-grid40
python tools/ncu_step.py > /dev/null 2>&1; ncu --set full --clock-control none --import-source on -k regex:bn_ --launch-count 4 -f -o gpurun_out/r2_bn python tools/ncu_step.py > gpurun_out/r2_ncu_bn.log 2>&1; tail -2 gpurun_out/r2_ncu_bn.log
