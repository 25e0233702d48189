# 8-GPU weak-scaling check: one chain per GPU (torchrun), both arms.
mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 30 > gpurun_out/scale_n$n.json 2>gpurun_out/scale_n$n.err; python -c "import json; d=json.load(open('gpurun_out/scale_n$n.json')); print($n, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
done
python bench.py --steps 30 --no-cpu-baseline > gpurun_out/scale_n1.json 2>gpurun_out/scale_n1.err; python -c "import json; d=json.load(open('gpurun_out/scale_n1.json')); print(1, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 8 --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-180
timeout 200 python scripts/small_case.py 2>&1 | tail -14
