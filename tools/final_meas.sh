set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_depth|k_seg_walk|k_seg_class|k_seg_scan|k_seg_place|k_march_blocks|k_commit_tma" -s 7 -c 7 -o gpurun_out/r02_integrate python tools/prof_march.py S2 > gpurun_out/prof_integrate.log 2>&1
python tools/bench_paths.py > gpurun_out/r02_paths.md 2> gpurun_out/r02_paths.err
tail -c 600 gpurun_out/r02_bench_1gpu.json
