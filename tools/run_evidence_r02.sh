set -x
cd $GRAFT_REPO_ROOT
B="python bench.py --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-fp32-frames --no-store-e2e --profile-passes 1 --no-graph-profile"
timeout 300 python tools/sweep_membound.py > gpurun_out/r02b_membound_sweep.md 2> gpurun_out/r02b_sweep.err
HULC2_RNN_COOP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/r02b_launches.csv $B --steps 3 > gpurun_out/r02b_launches.log 2>&1
HULC2_RNN_COOP=0 timeout 900 ncu --set full --clock-control none -k regex:"rnn_cluster2|ssm_reg|adam_kernel|logistic_loss_kernel|kl_fwd|kl_bwd|infonce|attention|colsum_wide|transpose_small|add_pos_bwd|frames_u8_pack" -c 40 -f -o gpurun_out/r02b_misc $B --steps 1 > gpurun_out/r02b_misc.log 2>&1
ncu -i gpurun_out/r02b_misc.ncu-rep --page raw --csv > gpurun_out/r02b_misc_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_misc.ncu-rep
HULC2_RNN_COOP=0 timeout 1200 ncu --set full --clock-control none -k regex:"gemm_tma" -c 96 -f -o gpurun_out/r02b_gemm $B --steps 1 > gpurun_out/r02b_gemm.log 2>&1
ncu -i gpurun_out/r02b_gemm.ncu-rep --page raw --csv > gpurun_out/r02b_gemm_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_gemm.ncu-rep
HULC2_SWEEP_REPS=1 HULC2_SWEEP_ONLY=large timeout 900 ncu --set full --clock-control none -k regex:"logistic_loss_kernel|logistic_sample|kl_fwd_grid|kl_bwd|layernorm|ssm_reg|gru_cell|lstm_cell|frame_kernel" -f -o gpurun_out/r02b_sweep_large python tools/sweep_membound.py > gpurun_out/r02b_sweep_large.log 2>&1
ncu -i gpurun_out/r02b_sweep_large.ncu-rep --page raw --csv > gpurun_out/r02b_sweep_large_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_sweep_large.ncu-rep
ls -la gpurun_out | tail -15
