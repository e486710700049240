mkdir -p gpurun_out
{ nproc; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; } > gpurun_out/topo_n8.txt 2>&1
for B in 1 0; do
LVN_NUMA_BIND=$B python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$B bench.py --gpus 8 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8_bind$B.json 2> gpurun_out/bench_n8_bind$B.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_bind$B.json').read().strip().splitlines()[-1]); print('bind$B', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d.get('host_numa'))"
done
for N in 8 4; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N profiles/bench_update_multi.py 2>gpurun_out/update_multi_n$N.err | tee gpurun_out/update_multi_n$N.json | tail -1; done
