for f in "" "--no-overlap" "--no-shard" "--no-overlap --no-shard"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 --no-profile $f 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
