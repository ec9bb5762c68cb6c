set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5_smoke.log 2>&1; tail -1 gpurun_out/r5_smoke.log
timeout 900 python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; tail -c 600 gpurun_out/r5_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5_ref.json 2> gpurun_out/r5_ref.err; tail -c 300 gpurun_out/r5_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r5_ncu1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5_launches_c3.csv python bench.py --workload config3 --frames 8 --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r5_ncu2.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r5_c2 python scripts/prof_target.py > gpurun_out/r5_ncu3.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r5_aa python scripts/prof_target.py --workload config3 --frames 1 > gpurun_out/r5_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
