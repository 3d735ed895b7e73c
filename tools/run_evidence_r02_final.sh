# Final r02 evidence on ONE box (1 x B200): parity suite, smoke, the default bench line, bench lines of the other configs, the ncu launch
# list and --set full captures of the kernels that changed in the last third of the round.  Outputs under gpurun_out/r02f_*.
set -x
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02f_pytest.log 2>&1; tail -3 $O/r02f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02f_smoke.log 2>&1; tail -2 $O/r02f_smoke.log
timeout 600 python bench.py --dump-profile $O/r02f_percall.json > $O/r02f_bench.json 2> $O/r02f_bench.err; tail -c 300 $O/r02f_bench.json
N="python bench.py --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-fp32-frames --no-store-e2e --profile-passes 1 --no-graph-profile"
HULC2_RNN_COOP=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4300 --csv --log-file $O/r02f_launches.csv $N --steps 3 > $O/r02f_launches.log 2>&1
HULC2_RNN_COOP=0 timeout 420 ncu --set full --clock-control none -k regex:"gemm_tma|attention|rnn_cluster2|conv_halo_kernel" -c 110 -f -o $O/r02f_full $N --steps 1 > $O/r02f_full.log 2>&1
ncu -i $O/r02f_full.ncu-rep --page raw --csv > $O/r02f_full_raw.csv 2>/dev/null
rm -f $O/r02f_full.ncu-rep
ls -la $O | grep r02f
