# 2-GPU pass: torchrun bench (weak scaling, one chain per GPU), evolve + multievolve end to end on the bundled example.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python bench.py --no-cpu-baseline > gpurun_out/scale_n1.json 2>gpurun_out/scale_n1.err; cut -c1-160 gpurun_out/scale_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/scale_n2.json 2>gpurun_out/scale_n2.err; cut -c1-160 gpurun_out/scale_n2.json; tail -3 gpurun_out/scale_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-200
rm -rf /tmp/run1 && mkdir -p /tmp/run1
( time python -m phylowgs_b200.host.evolve -O /tmp/run1/single -B 20 -s 30 -r 5 tests/golden/ssm_data.txt tests/golden/cnv_data.txt ) 2>&1 | tail -8
ls -la /tmp/run1/single; tail -3 /tmp/run1/single/mcmc_samples.txt
( time python -m phylowgs_b200.host.multievolve -n 2 -r 1 2 -O /tmp/run1/chains --ssms tests/golden/ssm_data.txt --cnvs tests/golden/cnv_data.txt -B 10 -s 20 ) 2>&1 | tail -12
python - <<'PY'
import zipfile
z = zipfile.ZipFile('/tmp/run1/chains/trees.zip'); n = z.namelist(); print(len(n), n[:3], n[-2:])
PY
