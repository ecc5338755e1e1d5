import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench

if __name__ == "__main__":
    for tpw, nimg in [(2, 64), (4, 64), (8, 32), (16, 16)]:
        t0 = time.time()
        ips, dt, split = bench.cpu_pool_images_per_sec("dino_vits16", 480, 5, nimg, 1, 1, threads_per_worker=tpw)
        print(f"threads/worker {tpw}: {ips:.2f} img/s  split {split}  (wall {time.time()-t0:.0f}s)", flush=True)
