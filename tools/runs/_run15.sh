python tools/gemm_anchor.py > gpurun_out/r2_gemm_anchor.txt 2>&1; cat gpurun_out/r2_gemm_anchor.txt | tail -9
KF='regex:cnn_fwd_kernel|mlp_chain_kernel|hips_fsa_direct_kernel|cnn_bwd_all_kernel|bn_fwd_kernel|bn_bwd_kernel|gemm_tf32_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 15 --launch-count 9 -f -o gpurun_out/r2_step python tools/ncu_step.py > gpurun_out/r2_ncu_step.log 2>&1; tail -3 gpurun_out/r2_ncu_step.log
ls -la gpurun_out/r2_step.ncu-rep
python tools/kernel_times.py > gpurun_out/r2_ktimes5.txt 2>&1; head -8 gpurun_out/r2_ktimes5.txt
