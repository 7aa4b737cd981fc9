# Round-2 (final) evidence on one B200; everything lands in gpurun_out/ under the prefix r02f_.
#   gpurun --timeout 1500 -- 'bash tools/run_round_evidence.sh'
set -x
P=gpurun_out/r02f
timeout 600 python -m pytest tests -m gpu -x -q > ${P}_pytest.log 2>&1; echo pytest rc=$?; tail -2 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee ${P}_smoke.log
python bench.py > ${P}_bench.json 2> ${P}_bench.err; tail -c 600 ${P}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2>/dev/null
python tools/bench_configs_multi.py > ${P}_configs_1gpu.json 2>/dev/null
# launch list of a short bench run (after the same command has exited 0 without ncu)
python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > ${P}_ncu_list.log 2>&1
# full captures: the persistent INT8 kernel (k = 20), and at k = 40 the contract-only kernel and the Cholesky kernel
TAG=plain python tools/i8_time.py > ${P}_i8time.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dla_loglik_i8p -s 2 -c 1 -o ${P}_i8p python tools/i8_time.py > ${P}_ncu_i8p.log 2>&1
K=40 QB=37 TAG=plain python tools/i8_time.py >> ${P}_i8time.log 2>&1 && K=40 QB=37 REPS=2 ncu --set full --clock-control none --import-source on -k regex:gram_contract -s 1 -c 1 -o ${P}_contract python tools/i8_time.py > ${P}_ncu_contract.log 2>&1
K=40 QB=37 REPS=2 ncu --set full --clock-control none -k regex:cholesky_kernel -s 1 -c 1 -o ${P}_chol python tools/i8_time.py > ${P}_ncu_chol.log 2>&1
K=40 QB=37 REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_k40_launches.csv python tools/i8_time.py > /dev/null 2>&1
cat ${P}_i8time.log
# summaries on the box (gpurun brings back at most 64 MiB: the three reports together exceed that)
python tools/ncu_summary.py ${P}_i8p.ncu-rep ${P}_ncu_loglik_i8_full.json > /dev/null
python tools/ncu_stalls.py ${P}_i8p.ncu-rep ${P}_ncu_i8p_stalls.txt > /dev/null
python tools/ncu_summary.py ${P}_contract.ncu-rep ${P}_ncu_contract_full.json > /dev/null
python tools/ncu_stalls.py ${P}_contract.ncu-rep ${P}_ncu_contract_stalls.txt > /dev/null
python tools/ncu_summary.py ${P}_chol.ncu-rep ${P}_ncu_cholesky_full.json > /dev/null
rm -f ${P}_contract.ncu-rep ${P}_chol.ncu-rep
ls -la gpurun_out/ | tail -30
