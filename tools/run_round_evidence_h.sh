# Refresh of the 1-GPU evidence after the lane-contiguous table / pixel planes (prefix r02h_); the contract-only and
# Cholesky kernels are unchanged since tools/run_round_evidence.sh (r02f_).
set -x
P=gpurun_out/r02h
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.log 2>&1; echo pytest rc=$?; tail -2 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee ${P}_smoke.log
python bench.py > ${P}_bench.json 2> ${P}_bench.err; tail -c 300 ${P}_bench.json
python tools/bench_configs_multi.py > ${P}_configs_1gpu.json 2>/dev/null
python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_list.log 2>&1
TAG=plain python tools/i8_time.py > ${P}_i8time.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dla_loglik_i8p -s 2 -c 1 -o ${P}_i8p python tools/i8_time.py > ${P}_ncu_i8p.log 2>&1
K=40 QB=37 TAG=plain python tools/i8_time.py >> ${P}_i8time.log 2>&1
cat ${P}_i8time.log
python tools/ncu_summary.py ${P}_i8p.ncu-rep ${P}_ncu_loglik_i8_full.json > /dev/null
python tools/ncu_stalls.py ${P}_i8p.ncu-rep ${P}_ncu_i8p_stalls.txt > /dev/null
ls -la gpurun_out/ | tail -16
