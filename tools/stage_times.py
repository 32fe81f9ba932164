#!/usr/bin/env python3
"""Development aid: per-stage CUDA-event times of the C2 bench workload under different BSB_* tuning variables.

python tools/stage_times.py [--batches 2] "VAR=1,VAR2=3" "VAR=2" ...   (an empty string = defaults)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, default=2)
    ap.add_argument('--genome-mb', type=int, default=250)
    ap.add_argument('--batch-pairs', type=int, default=266666)
    ap.add_argument('configs', nargs='*', default=[''])
    a = ap.parse_args()
    from bsbolt_b200 import _native, index_db
    work = os.path.join('/tmp/bsb_bench', f'g{a.genome_mb}')
    db = os.path.join(work, 'db', 'BSB_ref.fa')
    fa = bench.ensure_genome(work, os.path.dirname(db), a.genome_mb)
    jobs = bench.prepare_workload(work, fa, a.batches, a.batch_pairs)
    sims = bench.run_simulation(jobs, min(8, len(jobs)))
    f1 = os.path.join(work, 'st_1.fq'); f2 = os.path.join(work, 'st_2.fq')
    bench.concat([p[0] for p, n in sims], f1); bench.concat([p[1] for p, n in sims], f2)
    if not os.path.exists(db + '.sa'):
        index_db.build_database(fa, os.path.join(work, 'db'), device=0)
    null = os.open(os.devnull, os.O_WRONLY)
    argv = ['mem'] + bench.LAUNCHER_ARGS + ['-t', '1', '-K', str(a.batch_pairs * 300), '-v', '1', db, f1, f2]
    names = ('h2d', 'convert', 'seed', 'scan_sa', 'chain', 'extend', 'pestat', 'final')
    for cfg in a.configs:
        kv = [x.split('=') for x in cfg.split(',') if x]
        for k, v in kv:
            os.environ[k] = v
        idx = _native.Index(db, 0)        # some variables (index layout) are read when the index is loaded
        _native.mem_main(argv, index=idx, out_fd=null, log_fd=null)   # warm-up
        rc, st = _native.mem_main(argv, index=idx, out_fd=null, log_fd=null)
        hbm = idx.hbm_bytes
        idx.close()
        for k, v in kv:
            del os.environ[k]
        nb = max(1, st['n_batches'])
        stages = ' '.join(f'{n}={v / nb:.2f}' for n, v in zip(names, st['ms_stage']))
        print(f'[{cfg or "default"}] rc={rc} kernels={st["ms_kernels"] / nb:.2f} ms/batch | {stages} | select={st["ms_select"] / nb:.2f} tasks={st["ms_tasks"] / nb:.2f} | index {hbm / 1e9:.2f} GB', flush=True)


if __name__ == '__main__':
    main()
