#!/bin/bash
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "cfs_quota: $(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null) period $(cat /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null)"
cat /sys/fs/cgroup/cpu.stat 2>/dev/null
python -c "import os;print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))"
nproc; cat /proc/self/cgroup | head -5; ls /sys/fs/cgroup | head -30
cat /proc/loadavg
# raw CPU scaling check: N busy processes for 2 s, count iterations
python - <<'PY'
import multiprocessing as mp, time
def burn(q):
    t=time.perf_counter(); n=0
    while time.perf_counter()-t<1.5:
        for _ in range(10000): n+=1
    q.put(n)
if __name__=="__main__":
    for k in (1,8,16,32,64):
        q=mp.Queue(); ps=[mp.Process(target=burn,args=(q,)) for _ in range(k)]
        [p.start() for p in ps]; tot=sum(q.get() for _ in ps); [p.join() for p in ps]
        print("procs",k,"iters/s per proc", round(tot/k/1.5/1e6,2),"M total", round(tot/1.5/1e6,1),"M")
PY
cat /sys/fs/cgroup/cpu.stat 2>/dev/null
