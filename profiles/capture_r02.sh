set -x
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc $?"
for c in rollout beam step lone; do
  case $c in rollout) k=env_rollout_kernel; s=2;; beam) k=beam_search_kernel; s=2;; step) k=env_step_fused; s=7;; lone) k=team_games_kernel; s=0;; esac
  timeout 60 python profiles/ncu_cases.py $c > /dev/null 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/prof_${c}_r02 -f python profiles/ncu_cases.py $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c rc $?"
done
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:env_step_fused -s 7 -c 1 -o gpurun_out/prof_step_warm_r02 -f python profiles/ncu_cases.py step > gpurun_out/ncu_step_warm.log 2>&1; echo "ncu step warm rc $?"
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:env_step_fused -s 7 -c 1 -o gpurun_out/prof_step_plain_r02 -f python profiles/ncu_cases.py step_plain > gpurun_out/ncu_step_plain.log 2>&1; echo "ncu step plain rc $?"
# (a --set full capture of play_games_kernel at this size -- 0.44 s per pass, ~45 passes with source counters -- did not finish in 280 s; the team kernel does)
for k in team; do timeout 400 ncu --set full --clock-control none --import-source on -k regex:${k}_games_kernel -c 1 -o gpurun_out/prof_games_${k}_r02 -f python profiles/ncu_cases.py games10k > gpurun_out/ncu_games_$k.log 2>&1; echo "ncu games $k rc $?"; done
ls -la gpurun_out/*.ncu-rep
