set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01f_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r01f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err
python bench.py --gram-digits 5 --no-cpu-baseline > gpurun_out/r01f_bench_5digits.json 2>/dev/null
python bench.py --gram-digits -1 --no-cpu-baseline > gpurun_out/r01f_bench_f64.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_reference.json 2>/dev/null
python tools/bench_configs.py > gpurun_out/r01f_bench_configs.json 2>/dev/null
python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --quasars 592 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_ncu_list.log 2>&1
TAG=plain python tools/i8_time.py > gpurun_out/r01f_i8time.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dla_loglik_i8p -s 2 -c 1 -o gpurun_out/prof_r01f python tools/i8_time.py > gpurun_out/r01f_ncu_full.log 2>&1
ls -la gpurun_out/prof_r01f.ncu-rep
