set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_n1_reference.json 2> gpurun_out/r02_bench_n1_reference.err
timeout 300 python bench.py --arith fast --no-extra > gpurun_out/r02_bench_n1_fast.json 2>/dev/null
timeout 400 python bench.py --config c4a --no-extra > gpurun_out/r02_bench_c4a.json 2> gpurun_out/r02_bench_c4a.err
timeout 300 python bench.py --config c4a --arith fast --no-extra > gpurun_out/r02_bench_c4a_fast.json 2>/dev/null
timeout 120 python tools/lone_latency.py > gpurun_out/r02_lone_latency_by_option.txt 2>&1
timeout 120 python tools/sor_small_rate.py > gpurun_out/r02_sor_small_rates.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_c3.csv python tools/one_pair.py c3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/r02_launch_list.csv python bench.py --steps 2 --warmup 1 --batch 16 --batch-handles 2 --streams 2 --no-extra > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_front -s 4 -c 1 -o gpurun_out/r02_ncu_front python tools/one_pair.py c3 > /dev/null 2>&1
ls -la gpurun_out | tail -20
