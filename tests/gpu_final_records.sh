python bench.py > gpurun_out/r02c_bench_bf16.json 2> gpurun_out/r02c_bench_bf16.err
python tests/gpu_train_timeline.py > gpurun_out/r02c_train_timeline.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_train_step_launches.csv python tests/gpu_train_step_target.py ncu > /dev/null 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02c_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_mlp|k_wgrad" -c 6 --profile-from-start off -o gpurun_out/r02c_train_kernels python tests/gpu_train_step_target.py ncu > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_bf16.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["clocks"])
s=d["train"]["device_side_step"]; print(s["iters_per_s"], s["ms_per_iter"], s["first_30_iters"], s["frac_of_tensor_roofline"])
print(d["train"]["iters_per_s"], d["parity_mode"]["value"], d["cpu_baseline"]["value"])
PY
head -1 gpurun_out/r02c_train_timeline.txt; tail -1 gpurun_out/r02c_pytest_gpu.log
