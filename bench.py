#!/usr/bin/env python3
"""bench.py -- headline benchmark of the `bsbolt Align` hot path on B200.

Metric (BASELINE.json): paired-end 150 bp WGBS reads aligned per second. Workload = configs[1]:
250 Mb synthetic genome (10 x 25 Mb), directional PE150 reads from the seeded simulator
(bsbolt_b200/simulate.py, conventions of `bsbolt Simulate`), `bsbolt Align` argv (Launcher.py:77-98)
with a fixed batch size -K. One "step" = one batch of --batch-pairs read pairs.

  value  reads/s with every timed batch already resident in HBM when the clock starts (BSB_RESIDENT_BENCH: all batches
         are parsed and uploaded first, then released to the device at once): reads / time until the last batch has
         left the device, results copied back; three batches in flight per GPU, exactly as in the product run
         (max over ranks)
  e2e    reads/s through the public API (bsb_mem_main: FASTQ files on the host -> SAM text to /dev/null),
         host parsing, H2D, kernels, D2H, SAM formatting all inside the timed region
  roofline   dominant kernel (SMEM seeding) against the measured HBM copy bandwidth
  cpu_baseline / --impl reference   the reference's own multithreaded CPU aligner (oracle/_ref/bwa, built
         from /root/reference by oracle/Makefile) on a bounded sample of the same reads, same argv

python bench.py --gpus N --steps K --warmup W [--impl reference]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAUNCHER_ARGS = ('-Y -A 1 -B 4 -D 0.5 -E 1,1 -L 30,30 -T 10 -U 17 -W 0 -c 500 -d 100 -k 19 -m 50 -r 1.5 -w 100 -y 20 '
                 '-O 6,6 -h 100,200 -e 0.1 -l 0.5 -n 5 -Z 0.95').split()
# algorithmic FM-index bytes per PE150 read on an out-of-cache index (reference layout)
SEED_BYTES_PER_READ = 66.3e3
SA_BYTES_PER_READ = 45.7e3


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def _simulate_step(args):
    fa, prefix, n_pairs, seed, first_id = args
    from bsbolt_b200 import simulate
    names, ctg = simulate.read_fasta(fa)
    paths, n = simulate.simulate_reads(names, ctg, prefix, n_pairs, seed=seed, truth=False, first_id=first_id,
                                       corrupt_frac=float(os.environ.get('BSB_SIM_CORRUPT', '0')), undirectional=bool(os.environ.get('BSB_SIM_UNDIRECTIONAL')))
    return paths, n


def prepare_workload(work, genome_mb, n_steps, batch_pairs, rank, seed0=1000, dist=None):
    """Genome + per-step FASTQ pairs (rank-specific reads, shared genome written by rank 0 only)."""
    from bsbolt_b200 import simulate
    os.makedirs(work, exist_ok=True)
    fa = os.path.join(work, 'genome.fa')
    if rank == 0 and not os.path.exists(fa + '.done'):
        n_ctg = 10
        simulate.make_genome(fa, [genome_mb * 1000000 // n_ctg] * n_ctg, seed=20240517)
        open(fa + '.done', 'w').write('ok')
    if dist:
        dist.barrier()
    jobs = []
    for s in range(n_steps):
        prefix = os.path.join(work, f'r{rank}_s{s}')
        jobs.append((fa, prefix, batch_pairs, seed0 + 7919 * rank + s, s * batch_pairs * 2))
    return fa, jobs


def run_simulation(jobs, workers):
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_simulate_step, jobs))


def concat(paths, dst, remove=False):
    with open(dst, 'wb') as o:
        for p in paths:
            with open(p, 'rb') as f:
                shutil.copyfileobj(f, o, 1 << 24)
            if remove:      # the per-step files are only needed once (keeps the scratch footprint of an 8-rank run down)
                os.remove(p)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm)}


def head_records(src, dst, n):
    """first n FASTQ records of src"""
    with open(src, 'rb') as f, open(dst, 'wb') as o:
        for _ in range(4 * n):
            line = f.readline()
            if not line:
                break
            o.write(line)


def reference_cpu_run(bwa, argv_tail, n_reads, threads, sam_out=os.devnull):
    """Times the reference aligner; the clock starts when the first batch has been read (index load excluded)."""
    cmd = [bwa, 'mem'] + LAUNCHER_ARGS + ['-t', str(threads)] + argv_tail
    t_first = None
    with open(sam_out, 'w') as null:
        p = subprocess.Popen(cmd, stdout=null, stderr=subprocess.PIPE, text=True)
        for line in p.stderr:
            if t_first is None and line.startswith('[M::process] read'):
                t_first = time.time()
        p.wait()
    t_end = time.time()
    if p.returncode != 0 or t_first is None:
        raise RuntimeError('reference aligner failed')
    return n_reads / (t_end - t_first), t_end - t_first


def compare_sam(ref_path, my_path):
    """record-by-record identity of two SAM files (headers: all but @PG, whose CL differs by construction)"""
    n = same = 0
    hdr_same = True
    with open(ref_path) as fa, open(my_path) as fb:
        la = [l for l in fa if not l.startswith('@PG')]
        lb = [l for l in fb if not l.startswith('@PG')]
    ha = [l for l in la if l.startswith('@')]; hb = [l for l in lb if l.startswith('@')]
    hdr_same = ha == hb
    ra = la[len(ha):]; rb = lb[len(hb):]
    n = max(len(ra), len(rb))
    same = sum(1 for x, y in zip(ra, rb) if x == y)
    return {'sam_records': n, 'identical': same, 'identical_frac': same / n if n else None, 'header_identical': hdr_same}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=32)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--genome-mb', type=int, default=250)
    ap.add_argument('--batch-pairs', type=int, default=266666, help='read pairs per step (= -K 80 Mbp)')
    ap.add_argument('--cpu-sample-pairs', type=int, default=100000)
    ap.add_argument('--work', default=os.environ.get('BSB_BENCH_WORK', '/tmp/bsb_bench'))
    ap.add_argument('--keep', action='store_true')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    W, K = max(a.warmup, 0), max(a.steps, 1)
    dist = None
    if world > 1 and a.impl == 'b200':
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group('nccl')
        dist = dist_
    if a.impl == 'reference' and rank != 0:
        return 0
    cores = os.cpu_count() or 1
    work = os.path.join(a.work, f'g{a.genome_mb}')
    K_bases = a.batch_pairs * 300
    config = {'workload': f'WGBS PE150 directional, {a.genome_mb} Mb synthetic genome (10 contigs), seeded simulator; '
                          f'{a.batch_pairs} pairs per step (-K {K_bases}); index + batch working set >> 126 MB L2 (no flush needed)',
              'genome_mb': a.genome_mb, 'read_len': 150, 'paired': True, 'batch_pairs': a.batch_pairs,
              'align_args': ' '.join(LAUNCHER_ARGS + ['-K', str(K_bases)])}

    # ---------------- workload ----------------
    t0 = time.time()
    fa, jobs = prepare_workload(work, a.genome_mb, W + K, a.batch_pairs, rank, dist=dist)
    sim_workers = max(1, min(len(jobs), cores // max(world, 1), 8))
    sims = run_simulation(jobs, sim_workers)
    t_sim = time.time() - t0
    warm = [p for (paths, n) in sims[:W] for p in paths]
    timed_pairs = sum(n for (_, n) in sims[W:])
    f1 = os.path.join(work, f'r{rank}_timed_1.fq'); f2 = os.path.join(work, f'r{rank}_timed_2.fq')
    w1 = os.path.join(work, f'r{rank}_warm_1.fq'); w2 = os.path.join(work, f'r{rank}_warm_2.fq')
    if a.impl != 'reference':             # (the reference arm samples every step's own file)
        concat([paths[0] for (paths, n) in sims[W:]], f1, remove=True)
        concat([paths[1] for (paths, n) in sims[W:]], f2, remove=True)
        if W:
            concat(warm[0::2], w1, remove=True); concat(warm[1::2], w2, remove=True)

    def cleanup():
        if not a.keep:
            for f in [f1, f2, w1, w2] + [p for (paths, n) in sims for p in paths]:
                if os.path.exists(f):
                    os.remove(f)
    n_reads_timed = 2 * timed_pairs

    bwa = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
    db = os.path.join(work, 'db', 'BSB_ref.fa')

    if a.impl == 'reference':
        # the reference arm: the reference's own CPU implementation on all host threads, bounded sample per step
        if not os.path.exists(bwa):
            print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref/bwa is not built (no /root/reference here and no prebuilt copy)'}))
            return 0
        if not os.path.exists(db + '.bwt'):
            # the reference needs an index; it is built by the product's GPU builder when a device exists,
            # otherwise by the reference indexer itself
            try:
                from bsbolt_b200 import index_db
                index_db.build_database(fa, os.path.join(work, 'db'))
            except Exception:
                os.makedirs(os.path.join(work, 'db'), exist_ok=True)
                shutil.copy(fa, db)
                sh([bwa, 'index', '-a', 'bwtsw', db])
        # bounded sample per step: about two minutes of host-core work over the whole --steps/--warmup run
        sp = min(a.cpu_sample_pairs, a.batch_pairs, max(10000, 4000000 // (W + K)))
        rates, secs = [], []
        for s in range(W + K):
            paths = sims[s][0]
            s1 = os.path.join(work, 'ref_s1.fq'); s2 = os.path.join(work, 'ref_s2.fq')
            for src, dst in zip(paths, (s1, s2)):
                head_records(src, dst, sp)
            r, sec = reference_cpu_run(bwa, ['-K', str(K_bases), db, s1, s2], 2 * sp, cores)
            if s >= W:
                rates.append(r); secs.append(sec)
        total_reads = 2 * sp * K
        val = total_reads / sum(secs)
        line = {'metric': 'paired-end 150bp WGBS reads aligned/sec', 'value': val, 'unit': 'reads/s', 'n_gpus': a.gpus, 'steps': K,
                'warmup': W, 'ms_per_step': 1000 * sum(secs) / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'int32', 'data': 'synthetic', 'impl': 'reference', 'config': config,
                'cpu_baseline': {'value': val, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference',
                                 'sample': f'{sp} pairs per step of the same simulated reads, bwa mem -t {cores}, clock from first batch read to exit'},
                'e2e': {'value': val, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        cleanup()
        return 0

    # ---------------- B200 arm ----------------
    from bsbolt_b200 import _native, index_db
    device = local_rank if world > 1 else 0
    t0 = time.time()
    if rank == 0 and not os.path.exists(db + '.sa'):
        index_db.build_database(fa, os.path.join(work, 'db'), device=device)
    if dist:
        dist.barrier()
    t_index = time.time() - t0
    idx = _native.Index(db, device)
    argv_common = ['mem'] + LAUNCHER_ARGS + ['-t', '1', '-K', str(K_bases), '-v', '1']
    null = os.open(os.devnull, os.O_WRONLY)

    def run(fq1, fq2, env=None):
        for k, v in (env or {}).items():
            os.environ[k] = v
        try:
            t = time.time()
            rc, st = _native.mem_main(argv_common + [db, fq1, fq2], index=idx, out_fd=null, log_fd=null)
        finally:
            for k in (env or {}):
                del os.environ[k]
        if rc:
            raise RuntimeError(_native.last_error())
        return time.time() - t, st
    launches0 = 0
    if W:
        launches0 = run(w1, w2)[1]['kernel_launches']
    import torch  # only for the device synchronisation / rank reduction the bench contract asks for

    def fenced(fn):
        if dist:
            dist.barrier()
        torch.cuda.synchronize(device)
        r = fn()
        torch.cuda.synchronize(device)
        if dist:
            dist.barrier()
        return r
    sampler = ClockSampler(device)
    sampler.start()
    wall, st = fenced(lambda: run(f1, f2))                                                  # e2e: host files -> SAM
    _, st_res = fenced(lambda: run(f1, f2, {'BSB_RESIDENT_BENCH': '1'}))                     # value: inputs resident
    clocks = sampler.stop()
    _, st_one = fenced(lambda: run(f1, f2, {'BSB_RESIDENT_BENCH': '1', 'BSB_GPU_SLOTS': '1'}))  # clean per-kernel times
    # the reference's default output (`-O prefix`: bwa mem | stream_bam): FASTQ files -> BAM file, on the first 4 steps' reads
    bam_info = None
    if world == 1:
        nb = min(4, K) * a.batch_pairs
        b1 = os.path.join(work, 'bam_1.fq'); b2 = os.path.join(work, 'bam_2.fq'); bam_path = os.path.join(work, 'bench_out.bam')
        head_records(f1, b1, nb); head_records(f2, b2, nb)
        for level in (-1, 1):
            t = time.time()
            rc, st_b = _native.mem_main_bam(argv_common + [db, b1, b2], bam_path, index=idx, threads=0, level=level, log_fd=null)
            dt = time.time() - t
            if rc:
                raise RuntimeError(_native.last_error())
            bam_info = bam_info or {'api': 'bsb_mem_main_bam (FASTQ files on host -> BGZF/BAM file)', 'reads': 2 * nb, 'unit': 'reads/s'}
            bam_info['zlib_default' if level < 0 else f'zlib_level_{level}'] = {'value': 2 * nb / dt, 'wall_s': dt, 'file_bytes': os.path.getsize(bam_path)}
        os.remove(bam_path)
    ms_resident, ms_total_wall = st_res['sec_resident'] * 1000, wall * 1000
    ms_one = st_one['sec_resident'] * 1000
    reads_all = n_reads_timed
    per_rank = None
    if dist:
        mine = torch.tensor([ms_resident, ms_total_wall, ms_one], device=f'cuda:{device}', dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank = {'ms_resident': [round(float(x[0]), 1) for x in every], 'ms_e2e_wall': [round(float(x[1]), 1) for x in every]}
        t = torch.tensor([ms_resident, ms_total_wall, ms_one], device=f'cuda:{device}', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_total_wall, ms_one = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([n_reads_timed], device=f'cuda:{device}', dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        reads_all = float(c[0])
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        cleanup()
        return 0
    n_batches = max(1, st['n_batches'])
    value = reads_all / (ms_resident / 1000)
    e2e = reads_all / (ms_total_wall / 1000)
    seed_ms = st_one['ms_stage'][2] / n_batches
    reads_per_launch = n_reads_timed / n_batches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = SEED_BYTES_PER_READ * reads_per_launch / (seed_ms / 1000) / 1e9
    traffic, layout = None, {}
    try:   # one ncu --set full capture of the same kernel on the same workload (profiles/r01_ncu_final.md)
        t = json.load(open(os.path.join(ROOT, 'profiles', 'r01_seed_traffic.json')))
        traffic = t['dram_bytes_read'] + t['dram_bytes_write']
        layout = {'counted_extensions_per_launch': t['extensions_per_launch'],
                  'counted_bytes_reference_layout': 64 * (t['extensions_per_launch'] + t['two_block_64B']),
                  'counted_bytes_this_layout': 32 * (t['extensions_per_launch'] + t['two_block_32B'])}
    except Exception:
        pass
    cpu = None
    parity = None
    if os.path.exists(bwa) and world == 1:
        sp = min(a.cpu_sample_pairs, a.batch_pairs)
        s1 = os.path.join(work, 'cpu_s1.fq'); s2 = os.path.join(work, 'cpu_s2.fq')
        ref_sam = os.path.join(work, 'cpu_ref.sam'); my_sam = os.path.join(work, 'cpu_mine.sam')
        for src, dst in ((f1, s1), (f2, s2)):
            head_records(src, dst, sp)
        try:
            r, sec = reference_cpu_run(bwa, ['-K', str(K_bases), db, s1, s2], 2 * sp, cores, sam_out=ref_sam)
            cpu = {'value': r, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference',
                   'sample': f'first {sp} pairs of the timed reads, oracle/_ref/bwa mem -t {cores}, {sec:.1f} s, clock from first batch read to exit'}
            # the same sample through the product (untimed): record-level identity with the reference's SAM
            fd = os.open(my_sam, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            rc, _ = _native.mem_main(argv_common + [db, s1, s2], index=idx, out_fd=fd, log_fd=null)
            os.close(fd)
            if rc == 0:
                parity = compare_sam(ref_sam, my_sam)
        except Exception as e:  # noqa
            cpu = cpu or {'value': None, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference', 'sample': f'failed: {e}'}
    line = {'metric': 'paired-end 150bp WGBS reads aligned/sec', 'value': value, 'unit': 'reads/s', 'n_gpus': a.gpus, 'steps': n_batches,
            'warmup': W, 'ms_per_step': ms_resident / n_batches, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32', 'data': 'synthetic', 'config': config, 'clocks': clocks,
            'e2e': {'value': e2e, 'unit': 'reads/s', 'h2d_bytes_per_step': st['h2d_bytes'] // n_batches, 'd2h_bytes_per_step': st['d2h_bytes'] // n_batches,
                    'api': 'bsb_mem_main (FASTQ files on host -> SAM text to /dev/null)', 'wall_s': wall},
            'gpu_launches': st['kernel_launches'] - launches0, 'batches_in_flight_per_gpu': int(os.environ.get('BSB_GPU_SLOTS', '3')),
            'value_one_batch_in_flight': reads_all / (ms_one / 1000),
            'roofline': {'bound': 'hbm', 'kernel': 'k_seed3 (+ k_pack4, k_seed3_finish: SMEM seeding stage)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)', 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
                         'algorithmic_bytes_per_read': SEED_BYTES_PER_READ, 'kernel_ms_per_launch': seed_ms,
                         'reads_per_launch': reads_per_launch, **layout},
            'cpu_baseline': cpu, 'parity_vs_reference': parity, 'e2e_bam': bam_info, 'per_rank': per_rank, 'host_cores': cores,
            'stage_ms_per_step': {k: v / n_batches for k, v in zip(('h2d', 'convert', 'seed', 'scan_sa', 'chain', 'extend', 'pestat', 'final'), st_one['ms_stage'])},
            'final_split_ms_per_step': {'select': st_one['ms_select'] / n_batches, 'tasks': st_one['ms_tasks'] / n_batches, 'n_tasks': st_one['n_tasks'] // n_batches},
            'stage_note': 'CUDA-event stage times of the run with ONE batch in flight (with several in flight the stages of different batches overlap)',
            'host_busy_s': {'read': st['sec_read'], 'format': st['sec_format'], 'write': st['sec_write'], 'gpu_thread': st['sec_align']},
            'setup_s': {'simulate': t_sim, 'index_build_or_wait': t_index}, 'index_hbm_bytes': idx.hbm_bytes}
    print(json.dumps(line))
    os.close(null)
    cleanup()
    return 0


if __name__ == '__main__':
    sys.exit(main())
