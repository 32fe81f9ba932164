#!/usr/bin/env python3
"""Development aid: end-to-end wall time of bsb_mem_main on the C2 bench workload under BSB_* variables.

python tools/e2e_probe.py [--batches 12] "VAR=1,VAR2=3" ...   (an empty string = defaults)"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, default=12)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--bam', action='store_true', help='write a BAM file (bsb_mem_main_bam, default level: device deflate) instead of SAM to /dev/null')
    ap.add_argument('--warm-batches', type=int, default=0, help='warm-up on the first N batches only (0: all)')
    ap.add_argument('configs', nargs='*', default=[''])
    a = ap.parse_args()
    from bsbolt_b200 import _native, index_db
    work = os.path.join('/tmp/bsb_bench', 'g250')
    db = os.path.join(work, 'db', 'BSB_ref.fa')
    fa = bench.ensure_genome(work, os.path.dirname(db), 250)
    jobs = bench.prepare_workload(work, fa, a.batches, 266666)
    sims = bench.run_simulation(jobs, min(max(8, (os.cpu_count() or 8) - 4), len(jobs)))
    f1 = os.path.join(work, 'st_1.fq'); f2 = os.path.join(work, 'st_2.fq')
    bench.concat([p[0] for p, n in sims], f1); bench.concat([p[1] for p, n in sims], f2)
    if not os.path.exists(db + '.sa'):
        index_db.build_database(fa, os.path.join(work, 'db'), device=0)
    null = os.open(os.devnull, os.O_WRONLY)
    argv = ['mem'] + bench.LAUNCHER_ARGS + ['-t', '1', '-K', str(266666 * 300), '-v', '1', db, f1, f2]
    idx = _native.MultiIndex(db, list(range(a.gpus))) if a.gpus > 1 else _native.Index(db, 0)
    mem = (lambda av: _native.mem_main_multi(av, idx, out_fd=null, log_fd=null)) if a.gpus > 1 else (lambda av: _native.mem_main(av, index=idx, out_fd=null, log_fd=null))
    if a.bam:
        bam_path = os.environ.get('BSB_PROBE_BAM_PATH', os.path.join(work, 'probe.bam'))
        mem = (lambda av: _native.mem_main_multi_bam(av, bam_path, idx, log_fd=null)) if a.gpus > 1 else (lambda av: _native.mem_main_bam(av, bam_path, index=idx, log_fd=null))
    if a.warm_batches:
        w1 = os.path.join(work, 'stw_1.fq'); w2 = os.path.join(work, 'stw_2.fq')
        bench.head_records(f1, w1, a.warm_batches * 266666); bench.head_records(f2, w2, a.warm_batches * 266666)
        mem(argv[:-2] + [w1, w2])
    else:
        mem(argv)   # warm-up
    for cfg in a.configs:
        kv = [x.split('=') for x in cfg.split(',') if x]
        for k, v in kv:
            os.environ[k] = v
        t = time.time()
        rc, st = mem(argv)
        dt = time.time() - t
        for k, v in kv:
            del os.environ[k]
        if st['sec_resident'] > 0:
            print(f'[{cfg}] resident: {st["sec_resident"]:.3f} s = {2 * 266666 * a.batches / st["sec_resident"] / 1e6:.2f} M reads/s', flush=True)
        print(f'[{cfg or "default"}] rc={rc} wall {dt:.3f} s = {2 * 266666 * a.batches / dt / 1e6:.2f} M reads/s | read {st["sec_read"]:.3f} format {st["sec_format"]:.3f} '
              f'gpu threads {st["sec_align"]:.3f} write {st["sec_write"]:.3f} | kernels {st["ms_kernels"] / max(1, st["n_batches"]):.1f} ms/batch h2d {st["ms_h2d"] / max(1, st["n_batches"]):.1f} d2h {st["ms_d2h"] / max(1, st["n_batches"]):.1f} '
              f'text {st["ms_text"] / max(1, st["n_batches"]):.2f} bam {st["ms_bam"] / max(1, st["n_batches"]):.2f} ({st["bam_raw_bytes"]} -> {st["bam_bgzf_bytes"]} bytes in {st["bam_blocks"]} blocks)', flush=True)


if __name__ == '__main__':
    main()
