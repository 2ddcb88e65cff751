for f in "" "--no-factored-sh"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 --no-profile $f 2>/dev/null | tail -1 > gpurun_out/scale_r2_v7_n8$f.json; python -c "import sys,json; d=json.loads(open('gpurun_out/scale_r2_v7_n8$f.json').read()); print('[$f]', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
done
