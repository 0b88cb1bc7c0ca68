cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_scan_topk -s 10 -c 2 -o gpurun_out/prof_k1_final python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_full_k1.log 2>&1
timeout 600 python bench.py --workload c3 --no-cpu > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
timeout 600 python bench.py --workload c4 --no-cpu > gpurun_out/final_bench_c4.json 2> gpurun_out/final_bench_c4.err
timeout 600 python bench.py --workload t10m --no-cpu > gpurun_out/final_bench_t10m.json 2> gpurun_out/final_bench_t10m.err
for f in c3 c4 t10m; do cut -c1-330 gpurun_out/final_bench_$f.json; done
ls -la gpurun_out/prof_k1_final.ncu-rep
