# ncu evidence of the final build (run on the GPU box from the repo root; summaries go to profiles/ via scripts/summarize_ncu.py)
set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r7_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r7_ncu1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r7_launches_c3.csv python bench.py --workload config3 --frames 8 --steps 2 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r7_ncu2.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r7_c2 python scripts/prof_target.py > gpurun_out/r7_ncu3.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/prof_r7_aa python scripts/prof_target.py --workload config3 --frames 1 > gpurun_out/r7_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_gemm -c 2 -f -o gpurun_out/prof_r7_gemm python scripts/time_blend_tc.py > gpurun_out/r7_ncu5.log 2>&1
ls -la gpurun_out/prof_r7_*.ncu-rep
