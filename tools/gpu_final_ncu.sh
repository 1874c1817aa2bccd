# ncu evidence of a round: --set full of the chain kernels (tools/profile_one.py, second pass) and the launch list of one bench step
set -x
TAG=${TAG:-r02e}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mlp_chain' -s 7 -c 7 -o gpurun_out/${TAG}_prof python tools/profile_one.py > gpurun_out/${TAG}_ncu_prof.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ls -la gpurun_out | grep ${TAG}_
