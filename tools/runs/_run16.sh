timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "cnn_direct or fused_step or lookahead" > gpurun_out/r2_t12.log 2>&1; tail -3 gpurun_out/r2_t12.log
python tools/step_timeline.py > gpurun_out/r2_timeline1g.txt 2>&1; tail -14 gpurun_out/r2_timeline1g.txt
python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_b16_n1.log 2>&1; tail -1 gpurun_out/r2_b16_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N1', d['ms_per_step'], d['ms_per_step_p10_p50_p90'], d['value'], d['e2e']['value'], d['exposed_push_pull_ms_per_step'], d['gpu_launches_per_step'])"
GEOMX_GEMM_STAGES=3 python tools/gemm_anchor.py > gpurun_out/r2_gemm_anchor_s3.txt 2>&1; tail -8 gpurun_out/r2_gemm_anchor_s3.txt
GEOMX_GEMM_STAGES=2 python tools/gemm_anchor.py 2>&1 | tail -8 | head -3
