pr() { tail -1 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['ms_per_step_p10_p50_p90'], d['value'], d['e2e']['value'], d['exposed_push_pull_ms_per_step'], d['gpu_launches_per_step'], d['protocol_errors'], (d.get('clocks') or {}).get('sm_mhz'))" 2>/dev/null || echo "$2 FAILED: $(tail -2 $1 | cut -c1-300)"; }
timeout 90 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 tools/fabric_check.py > gpurun_out/r2_fab8.log 2>&1; grep -E "FABRIC_CHECK|fabric backend|rank 0 iter|rank 7 iter|rank 0 graph" gpurun_out/r2_fab8.log | head -10
timeout 90 python bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r2_b20_n8.log 2>&1; pr gpurun_out/r2_b20_n8.log N8-channels
GEOMX_STEP_OVERLAP=0 timeout 90 python bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r2_b20_n8_ll.log 2>&1; pr gpurun_out/r2_b20_n8_ll.log N8-single-ll
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 90 python bench.py --gpus 4 --steps 200 --warmup 20 > gpurun_out/r2_b20_n4.log 2>&1) &
(CUDA_VISIBLE_DEVICES=4,5,6,7 GEOMX_STEP_OVERLAP=0 timeout 90 python bench.py --gpus 4 --steps 200 --warmup 20 > gpurun_out/r2_b20_n4_ll.log 2>&1) &
wait
pr gpurun_out/r2_b20_n4.log N4-channels; pr gpurun_out/r2_b20_n4_ll.log N4-single-ll
timeout 90 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 tools/step_timeline.py > gpurun_out/r2_timeline8.txt 2>&1; grep -v "^W0\|OMP\|\*\*\*" gpurun_out/r2_timeline8.txt | grep -A12 "^rank 0\|^rank 5" | head -28
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 90 python bench.py --gpus 4 --steps 100 --warmup 10 --config bsc > gpurun_out/r2_cfg4_bsc.log 2>&1) &
(CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 90 python bench.py --gpus 4 --steps 100 --warmup 10 --config mpq_dgt > gpurun_out/r2_cfg4_mpq.log 2>&1) &
wait
(CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 90 python bench.py --gpus 4 --steps 100 --warmup 10 --config hfa > gpurun_out/r2_cfg4_hfa.log 2>&1) &
(CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 90 python bench.py --gpus 4 --steps 100 --warmup 10 --config mixed_sync > gpurun_out/r2_cfg4_mixed.log 2>&1) &
wait
pr gpurun_out/r2_cfg4_bsc.log N4-bsc; pr gpurun_out/r2_cfg4_mpq.log N4-mpq_dgt; pr gpurun_out/r2_cfg4_hfa.log N4-hfa; pr gpurun_out/r2_cfg4_mixed.log N4-mixed_sync
