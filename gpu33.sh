python bench.py --impl reference > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
cat gpurun_out/r1_bench_reference.json | cut -c1-600
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
cat gpurun_out/r1_bench_n1.json
tail -3 gpurun_out/r1_bench_n1.err
