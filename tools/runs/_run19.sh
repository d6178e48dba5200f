timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/r2_t15.log 2>&1; tail -3 gpurun_out/r2_t15.log
pr() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['ms_per_step_p10_p50_p90'], d['value'], d['e2e']['value'], d['exposed_push_pull_ms_per_step'], d['gpu_launches_per_step'], d['protocol_errors'])"; }
python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_b19_n1.log 2>&1; pr gpurun_out/r2_b19_n1.log N1
python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b19_n2.log 2>&1; pr gpurun_out/r2_b19_n2.log N2
GEOMX_STEP_OVERLAP=0 python bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r2_b19_n2_ll.log 2>&1; pr gpurun_out/r2_b19_n2_ll.log N2-single-ll
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/step_timeline.py > gpurun_out/r2_timeline2g.txt 2>&1; grep -v "^W0\|OMP\|\*\*\*" gpurun_out/r2_timeline2g.txt | tail -30
python bench.py --gpus 2 --steps 50 --warmup 10 --script > gpurun_out/r2_b19_n2_script.log 2>&1; pr gpurun_out/r2_b19_n2_script.log N2-script
python bench.py --gpus 2 --steps 50 --warmup 10 --script --hybridize > gpurun_out/r2_b19_n2_script_hyb.log 2>&1; pr gpurun_out/r2_b19_n2_script_hyb.log N2-script-hybridize; tail -3 gpurun_out/r2_b19_n2_script_hyb.log | cut -c1-300
