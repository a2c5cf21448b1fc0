# ncu captures of the final round-2 build (run on the GPU box from the repo root; outputs under gpurun_out/)
set -x
cd ${GRAFT_REPO_ROOT:-.}
# (1) launch list of ONE step with DRAM bytes per launch (serialised, cold cache: compare shares; traffic per launch)
if [ "$1" != "full-only" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_launches.csv python bench.py --profile-one-step --warmup 3 > gpurun_out/prof1.log 2>&1
fi
# (2) --set full: the last ViT attention launch (23rd) and the first decoder attention launch (24th) of the step
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn_fwd_persistent_kernel --launch-skip 22 -c 2 -o gpurun_out/r02b_full_attn -f python bench.py --profile-one-step --warmup 3 > gpurun_out/prof2.log 2>&1
ls -la gpurun_out/
