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

Multi-GPU (--gpus N, or N ranks under torchrun): the timed run is the PRODUCT's own N-GPU run -- ONE process (rank 0),
one FASTQ pair holding N x steps batches, read and cut into batches once, batch b aligned on device b mod N against that
device's resident copy of the index, ONE SAM stream in input order (bsb_mem_main_multi). No collective is on the data path;
the other torchrun ranks only join the NCCL barriers that fence the timed regions (they wait on the rendezvous store, not
on a spinning stream, so that their cores stay free for rank 0's reader). A step = one batch per GPU; per-GPU work is fixed
as N grows (weak scaling). The N-GPU output of a bounded sample is compared byte for byte with the one-GPU output.

python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2|c5]
"""
import argparse
import hashlib
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


def prepare_workload(work, fa, n_batches, batch_pairs, seed0=1000):
    """one simulation job (one FASTQ pair) per batch"""
    os.makedirs(work, exist_ok=True)
    return [(fa, os.path.join(work, f'b{s}'), batch_pairs, seed0 + s, s * batch_pairs * 2) for s in range(n_batches)]


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


def md5_file(path):
    import hashlib
    h = hashlib.md5()
    with open(path, 'rb') as f:
        for blk in iter(lambda: f.read(1 << 24), b''):
            h.update(blk)
    return h.hexdigest()


def check_index_md5(fa, db, genome_mb):
    """The benchmark index is written by the product's GPU builder; tests/golden/bench_index_md5.json holds the md5 of
    what the REFERENCE indexer (`bwa index -a bwtsw`) wrote for the same (seeded) genome in the build container."""
    try:
        want = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'bench_index_md5.json')))
    except Exception:
        return 'no golden md5 file'
    if want.get('genome_mb') != genome_mb or md5_file(fa) != want['genome_fa_md5']:
        return 'not applicable (different genome)'
    bad = [e for e, h in want['index_md5'].items() if md5_file(f'{db}.{e}') != h]
    return 'identical to bwa index (6 files)' if not bad else 'DIFFERS from bwa index: ' + ','.join(bad)


class ClockSampler:
    """SM clocks / throttle reasons of every GPU of the run during the timed region (B200_PROFILING.md recipe), read
    through NVML every 50 ms on a thread (nvidia-smi -lms buffers its pipe output for seconds)."""

    def __init__(self, gpus):
        self.gpus, self.rows, self.stop_flag, self.th = list(gpus), [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(g) for g in self.gpus]
        except Exception:
            self.nv = None
            return
        self.th = threading.Thread(target=self._poll, daemon=True)
        self.th.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            for g, h in zip(self.gpus, self.handles):
                try:
                    self.rows.append((g, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                      nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
                except Exception:
                    pass
            time.sleep(0.05)

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join()
        if not self.nv:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'note': 'NVML not available'}
        nv = self.nv
        per = {}
        for g, mhz, _, _ in self.rows:
            per.setdefault(g, []).append(mhz)
        med = {str(g): sorted(v)[len(v) // 2] for g, v in per.items()}
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8), 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), 'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
        reasons = sorted({n for _, _, _, r in self.rows for n, bit in names.items() if r & bit})
        out = {'sm_mhz': min(med.values()) if med else None, 'sm_max_mhz': max([m for _, _, m, _ in self.rows] or [0]) or None, 'reasons': reasons, 'samples': len(self.rows)}
        if len(med) > 1:
            out['sm_mhz_per_gpu'] = med
        return out


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


def reference_cpu_bam_run(bwa, stream_bam, argv_tail, n_reads, threads, bam_out):
    """The reference's default output pipeline, `bwa mem ... | stream_bam -@ T -o out.bam` (bsbolt/Align/AlignReads.py:51-56), timed like
    reference_cpu_run; the clock stops when the BAM file is closed."""
    cmd = [bwa, 'mem'] + LAUNCHER_ARGS + ['-t', str(threads)] + argv_tail
    t_first = None
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=False)
    q = subprocess.Popen([stream_bam, '-@', str(threads), '-o', bam_out], stdin=p.stdout, stderr=subprocess.DEVNULL)
    p.stdout.close()
    for line in p.stderr:
        if t_first is None and line.startswith(b'[M::process] read'):
            t_first = time.time()
    p.wait(); q.wait()
    t_end = time.time()
    if p.returncode != 0 or q.returncode != 0 or t_first is None:
        raise RuntimeError('reference aligner | stream_bam failed')
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


def sam_digest(path):
    """(blake2 digest, bytes, records) of a SAM file without its @PG line"""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    n = recs = 0
    with open(path, 'rb') as f:
        for line in f:
            if line.startswith(b'@PG'):
                continue
            h.update(line); n += len(line); recs += line[:1] != b'@'
    return h.hexdigest(), n, recs


class Ranks:
    """torchrun ranks around a one-process multi-GPU run: rank 0 works, the others only meet it at the fences. A waiting rank
    blocks on the rendezvous store (a socket), then joins the NCCL barrier -- no core spins while rank 0's host pipeline runs."""

    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.dist, self.n = rank, world, None, 0
        if world > 1:
            import datetime
            import torch
            import torch.distributed as dist
            backend = os.environ.get('BSB_BENCH_BACKEND', 'nccl')   # gloo: the CPU test of this class (tests/test_sharding_gloo.py)
            if backend == 'nccl':
                torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, timeout=datetime.timedelta(hours=2))
            self.dist = dist
            self.store = dist.distributed_c10d._get_default_store()

    def fence(self):
        if not self.dist:
            return
        key = f'bsb_fence_{self.n}'
        self.n += 1
        if self.rank == 0:
            self.store.set(key, '1')
        else:
            import datetime
            self.store.wait([key], datetime.timedelta(hours=2))
        self.dist.barrier()

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


N_FENCES = 10  # fences rank 0 passes in the B200 arm (b200_arm: 2 x 4 timed regions + 2); the waiting ranks pass the same number


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=32)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c2', 'c5'], help='c2: directional PE150 (headline); c5: undirectional library, 5 %% corrupted mates (mate rescue)')
    ap.add_argument('--genome-mb', type=int, default=250)
    ap.add_argument('--batch-pairs', type=int, default=266666, help='read pairs per batch (= -K 80 Mbp)')
    ap.add_argument('--cpu-sample-pairs', type=int, default=100000)
    ap.add_argument('--work', default=os.environ.get('BSB_BENCH_WORK', '/tmp/bsb_bench'))
    ap.add_argument('--keep', action='store_true')
    a = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    W, K = max(a.warmup, 0), max(a.steps, 1)
    if a.impl == 'reference' and rank != 0:
        return 0
    n_dev = world if world > 1 else max(1, a.gpus)
    cores = os.cpu_count() or 1
    c5 = a.config == 'c5'
    if c5:
        os.environ['BSB_SIM_UNDIRECTIONAL'] = '1'
        os.environ.setdefault('BSB_SIM_CORRUPT', '0.05')
    work = os.path.join(a.work, f'g{a.genome_mb}{a.config}')
    K_bases = a.batch_pairs * 300
    extra_args = ['-z'] if c5 else []
    lib = 'undirectional (-UN), 5 % of the second mates corrupted (mate rescue)' if c5 else 'directional'
    config = {'workload': f'WGBS PE150 {lib}, {a.genome_mb} Mb synthetic genome (10 contigs), seeded simulator; '
                          f'{a.batch_pairs} pairs per batch (-K {K_bases}), one batch per GPU per step; index + batch working set >> 126 MB L2 (no flush needed)',
              'name': a.config, 'genome_mb': a.genome_mb, 'read_len': 150, 'paired': True, 'batch_pairs': a.batch_pairs,
              'align_args': ' '.join(LAUNCHER_ARGS + extra_args + ['-K', str(K_bases)])}
    bwa = os.path.join(ROOT, 'oracle', '_ref', 'bwa')
    db = os.path.join(a.work, f'g{a.genome_mb}', 'db', 'BSB_ref.fa')   # the index does not depend on the read library

    if a.impl == 'reference':
        return reference_arm(a, W, K, work, db, bwa, cores, config, K_bases, extra_args)

    ranks = Ranks(rank, world, local_rank)
    if rank != 0:           # rank 0 drives every GPU; this rank's cores stay free for its reader
        for _ in range(N_FENCES):
            ranks.fence()
        ranks.close()
        return 0
    if world > 1:
        os.environ['BSB_ALL_CORES'] = '1'   # the native pipeline would otherwise take 1/LOCAL_WORLD_SIZE of the cores
    try:
        return b200_arm(a, W, K, n_dev, work, db, bwa, cores, config, K_bases, extra_args, ranks)
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)         # the launcher ends the waiting ranks


def ensure_genome(work, db_dir, genome_mb):
    from bsbolt_b200 import simulate
    os.makedirs(db_dir, exist_ok=True)
    fa = os.path.join(os.path.dirname(db_dir), 'genome.fa')
    if not os.path.exists(fa + '.done'):
        n_ctg = 10
        simulate.make_genome(fa, [genome_mb * 1000000 // n_ctg] * n_ctg, seed=20240517)
        open(fa + '.done', 'w').write('ok')
    return fa


def reference_arm(a, W, K, work, db, bwa, cores, config, K_bases, extra_args):
    """The reference's own CPU implementation (oracle/_ref/bwa mem, all host threads) on a bounded sample of every step."""
    if not os.path.exists(bwa):
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref/bwa is not built (no /root/reference here and no prebuilt copy)'}))
        return 0
    fa = ensure_genome(work, os.path.dirname(db), a.genome_mb)
    jobs = prepare_workload(work, fa, W + K, a.batch_pairs)
    # bounded sample per step: about a minute of host-core work over the whole --steps/--warmup run
    sp = min(a.cpu_sample_pairs, a.batch_pairs, max(10000, 4000000 // (W + K)))
    jobs = [(f, pre, sp, seed, first) for (f, pre, n, seed, first) in jobs]     # the first sp pairs of each step's reads
    sims = run_simulation(jobs, max(1, min(len(jobs), cores - 1, 16)))
    index_note = None
    if not os.path.exists(db + '.sa'):
        # The reference needs an index. `bwa index` takes 10.5 min on this genome, so the files come from the GPU builder --
        # run as a separate executable (bsbolt_b200/bwa index), nothing of the product is loaded into this process -- and are
        # accepted only if their md5 equals what the reference indexer wrote for the same genome (tests/golden/bench_index_md5.json)
        from bsbolt_b200 import index_db
        index_db.write_database_fasta(fa, os.path.dirname(db))
        p = sh([os.path.join(ROOT, 'bsbolt_b200', 'bwa'), 'index', '-a', 'bwtsw', db])
        if p.returncode != 0:
            sh([bwa, 'index', '-a', 'bwtsw', db])
            index_note = 'built by the reference indexer (no GPU builder available)'
    index_note = index_note or check_index_md5(fa, db, a.genome_mb)

    def cat(steps, dst):
        for k in (0, 1):
            concat([sims[s][0][k] for s in steps], f'{dst}_{k + 1}.fq')
        return [f'{dst}_1.fq', f'{dst}_2.fq']
    tail = extra_args + ['-K', str(K_bases), db]
    if W:   # warm-up steps: page cache, index in memory
        reference_cpu_run(bwa, tail + cat(range(W), os.path.join(work, 'ref_warm')), 2 * sp * W, cores)
    # the K timed steps as ONE run, so that the reference's own pipeline (read || align || write, fastmap.c:352) overlaps as it does in production
    val, sec = reference_cpu_run(bwa, tail + cat(range(W, W + K), os.path.join(work, 'ref_timed')), 2 * sp * K, cores)
    line = {'metric': 'paired-end 150bp WGBS reads aligned/sec', 'value': val, 'unit': 'reads/s', 'n_gpus': a.gpus, 'steps': K,
            'warmup': W, 'ms_per_step': 1000 * sec / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32', 'data': 'synthetic', 'impl': 'reference', 'config': config,
            'cpu_baseline': {'value': val, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference',
                             'sample': f'{sp} pairs per step of the same simulated reads, {K} steps in one run of oracle/_ref/bwa mem -t {cores}, '
                                       f'{sec:.1f} s, clock from first batch read to exit'},
            'e2e': {'value': val, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'index': index_note}
    # the reference's DEFAULT output (`-O prefix`): the same timed reads through `| stream_bam` into a BAM file -- the counterpart of
    # the B200 arm's `e2e_bam`
    stream_bam = os.path.join(os.path.dirname(bwa), 'stream_bam')
    if os.path.exists(stream_bam):
        try:
            bam_out = os.path.join(work, 'ref_out.bam')
            vb, sb = reference_cpu_bam_run(bwa, stream_bam, tail + [os.path.join(work, 'ref_timed_1.fq'), os.path.join(work, 'ref_timed_2.fq')], 2 * sp * K, cores, bam_out)
            line['e2e_bam'] = {'value': vb, 'unit': 'reads/s', 'wall_s': sb, 'file_bytes': os.path.getsize(bam_out),
                               'pipeline': f'oracle/_ref/bwa mem -t {cores} | oracle/_ref/stream_bam -@ {cores} -o file (bsbolt/Align/AlignReads.py:51-56)'}
            os.remove(bam_out)
        except Exception as e:  # noqa
            line['e2e_bam'] = {'error': str(e)}
    print(json.dumps(line))
    if not a.keep:
        for f in [p for (paths, n) in sims for p in paths] + [os.path.join(work, f'ref_{w}_{k}.fq') for w in ('warm', 'timed') for k in (1, 2)]:
            if os.path.exists(f):
                os.remove(f)
    return 0


def b200_arm(a, W, K, n_dev, work, db, bwa, cores, config, K_bases, extra_args, ranks):
    from bsbolt_b200 import _native, index_db
    devices = list(range(n_dev))
    # ---------------- workload: ONE input of n_dev x (W + K) batches ----------------
    t0 = time.time()
    fa = ensure_genome(work, os.path.dirname(db), a.genome_mb)
    jobs = prepare_workload(work, fa, n_dev * (W + K), a.batch_pairs)
    sims = run_simulation(jobs, max(1, min(len(jobs), cores - 2, 28)))
    t_sim = time.time() - t0
    nw = n_dev * W
    timed_pairs = sum(n for (_, n) in sims[nw:])
    f1 = os.path.join(work, 'timed_1.fq'); f2 = os.path.join(work, 'timed_2.fq')
    w1 = os.path.join(work, 'warm_1.fq'); w2 = os.path.join(work, 'warm_2.fq')
    concat([paths[0] for (paths, n) in sims[nw:]], f1, remove=True)
    concat([paths[1] for (paths, n) in sims[nw:]], f2, remove=True)
    if W:
        concat([paths[0] for (paths, n) in sims[:nw]], w1, remove=True); concat([paths[1] for (paths, n) in sims[:nw]], w2, remove=True)
    scratch = [f1, f2, w1, w2]
    n_reads_timed = 2 * timed_pairs

    t0 = time.time()
    if not os.path.exists(db + '.sa'):
        index_db.build_database(fa, os.path.dirname(db), device=0)
    t_index = time.time() - t0
    index_note = check_index_md5(fa, db, a.genome_mb)
    idx = _native.MultiIndex(db, devices) if n_dev > 1 else _native.Index(db, 0)
    argv_common = ['mem'] + LAUNCHER_ARGS + extra_args + ['-t', '1', '-K', str(K_bases), '-v', '1']
    null = os.open(os.devnull, os.O_WRONLY)

    def mem(argv, out_fd):
        if n_dev > 1:
            return _native.mem_main_multi(argv, idx, out_fd=out_fd, log_fd=null)
        return _native.mem_main(argv, index=idx, out_fd=out_fd, log_fd=null)

    def run(fq1, fq2, env=None):
        for k, v in (env or {}).items():
            os.environ[k] = v
        try:
            t = time.time()
            rc, st = mem(argv_common + [db, fq1, fq2], null)
        finally:
            for k in (env or {}):
                del os.environ[k]
        if rc:
            raise RuntimeError(_native.last_error())
        return time.time() - t, st
    if W:
        # W untimed warm-up steps -- repeated until every batch context of every device (three per GPU) has seen two batches, so
        # that no device buffer is sized and no host buffer page-locked inside the timed region whatever W the caller chose
        for _ in range(max(1, -(-6 // W))):
            run(w1, w2)
    import torch  # only for the device synchronisation the bench contract asks for (and the NCCL fences under torchrun)

    def sync():
        for d in devices:
            torch.cuda.synchronize(d)

    def fenced(fn):
        ranks.fence(); sync()
        r = fn()
        sync(); ranks.fence()
        return r
    sampler = ClockSampler(devices)
    sampler.start()
    wall, st = fenced(lambda: run(f1, f2))                                                  # e2e: host files -> SAM
    _, st_res = fenced(lambda: run(f1, f2, {'BSB_RESIDENT_BENCH': '1'}))                     # value: inputs resident
    clocks = sampler.stop()
    _, st_one = fenced(lambda: run(f1, f2, {'BSB_RESIDENT_BENCH': '1', 'BSB_GPU_SLOTS': '1'}))  # clean per-kernel times
    ranks.fence()
    # ---- the N-GPU output against the one-GPU output of the same input (bounded sample, smaller batches) ----
    multi_identity = None
    if n_dev > 1:
        sp = 40000
        nb = 2 * n_dev
        m1 = os.path.join(work, 'mg_1.fq'); m2 = os.path.join(work, 'mg_2.fq')
        head_records(f1, m1, sp * nb); head_records(f2, m2, sp * nb)
        argv = ['mem'] + LAUNCHER_ARGS + extra_args + ['-t', '1', '-K', str(sp * 300), '-v', '1', db, m1, m2]
        outs = []
        for tag in ('one', 'multi'):
            path = os.path.join(work, f'mg_{tag}.sam')
            fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            if tag == 'one':
                rc, st_m = _native.mem_main(argv, index=idx.parts[0], out_fd=fd, log_fd=null)
            else:
                rc, st_m = _native.mem_main_multi(argv, idx, out_fd=fd, log_fd=null)
            os.close(fd)
            if rc:
                raise RuntimeError(_native.last_error())
            outs.append(sam_digest(path) + (st_m['n_batches'],))
            scratch.append(path)
        scratch += [m1, m2]
        multi_identity = {'batches': outs[1][3], 'sam_bytes': outs[1][1], 'records': outs[1][2],
                          'identical_to_one_gpu_output': outs[0][:3] == outs[1][:3]}
    # the reference's default output (`-O prefix`: bwa mem | stream_bam): FASTQ files -> BAM file.
    # Default level (what `bsbolt Align -O` uses): arbiter, BAM records and BGZF deflate on the device (bsb_bam.h, bsb_deflate.h), the
    # host appends the blocks to the file -- timed on the SAME input as `e2e`, fenced the same way. An explicit zlib level keeps the
    # round-1 path (SAM text to the host, encoded and deflated by the host cores): timed on the first 4 batches, where the two
    # files are also inflated (zlib, here) and their BAM streams compared.
    bam_path = os.path.join(work, 'bench_out.bam')

    def mem_bam(fq1, fq2, level, path=None):
        path = path or bam_path
        if path != os.devnull and os.path.exists(path):
            os.remove(path)                      # a fresh output file: dropping the previous run's cached pages is not part of the run
        t = time.time()
        if n_dev > 1:
            rc, st_b = _native.mem_main_multi_bam(argv_common + [db, fq1, fq2], path, idx, threads=0, level=level, log_fd=null)
        else:
            rc, st_b = _native.mem_main_bam(argv_common + [db, fq1, fq2], path, index=idx, threads=0, level=level, log_fd=null)
        if rc:
            raise RuntimeError(_native.last_error())
        return time.time() - t, st_b

    # from four GPUs on the timed input makes a BAM file of 10 GB and more: the blocks are discarded there (like `e2e` discards its
    # text) rather than risking the box's scratch space
    bam_sink = os.devnull if n_dev >= 4 else None

    def bam_timed():
        try:
            return mem_bam(f1, f2, -1, bam_sink)
        except Exception as e:  # noqa  (the SAM line above stands on its own; a failure here is reported, not fatal)
            return None, str(e)
    try:
        mem_bam(w1 if W else f1, w2 if W else f2, -1, bam_sink)                           # (sizes the BAM stage's buffers)
    except Exception:  # noqa
        pass
    dt, st_b = fenced(bam_timed)
    if dt is None:
        bam_info = {'error': st_b}
    else:
        nbt = max(1, st_b['n_batches'])
        bam_info = {'api': 'bsb_mem_main_bam%s (one FASTQ pair on the host -> one BGZF/BAM file), default level: compressed on the device' % ('_multi' if n_dev > 1 else ''),
                    'unit': 'reads/s', 'value': n_reads_timed / dt, 'reads': n_reads_timed, 'wall_s': dt, 'sink': 'file' if bam_sink is None else '/dev/null',
                    'file_bytes': os.path.getsize(bam_path) if bam_sink is None else None,
                    'd2h_bytes_per_step': st_b['d2h_bytes'] / nbt * n_dev, 'bam_bytes_per_step': st_b['bam_raw_bytes'] / nbt * n_dev,
                    'bgzf_bytes_per_step': st_b['bam_bgzf_bytes'] / nbt * n_dev,
                    'device_ms_per_batch': {'sam_text': st_b['ms_text'] / nbt, 'arbiter_records_deflate': st_b['ms_bam'] / nbt}}
    if dt is not None and bam_sink is None:
        try:   # the same run with the blocks discarded, like `e2e` discards its SAM text: what the file system costs above
            dt0, _ = mem_bam(f1, f2, -1, os.devnull)
            bam_info['to_dev_null'] = {'value': n_reads_timed / dt0, 'wall_s': dt0}
        except Exception as e:  # noqa
            bam_info['to_dev_null'] = {'error': str(e)}
    if n_dev == 1 and dt is not None:
        import gzip
        nb = min(4, K) * a.batch_pairs
        b1 = os.path.join(work, 'bam_1.fq'); b2 = os.path.join(work, 'bam_2.fq')
        head_records(f1, b1, nb); head_records(f2, b2, nb)
        sub = {}
        for name, level in (('device_deflate', -1), ('host_zlib_level_1', 1), ('host_zlib_level_6', 6)):
            dt, st_b = mem_bam(b1, b2, level)
            sub[name] = {'value': 2 * nb / dt, 'wall_s': dt, 'file_bytes': os.path.getsize(bam_path)}
            if level < 6:
                h = hashlib.sha256()
                with gzip.open(bam_path, 'rb') as f:
                    for chunk in iter(lambda: f.read(1 << 24), b''):
                        h.update(chunk)
                sub[name]['raw_sha256'] = h.hexdigest()
        sub['reads'] = 2 * nb
        sub['device_stream_identical_to_host_stream'] = sub['device_deflate']['raw_sha256'] == sub['host_zlib_level_1']['raw_sha256']
        bam_info['first_4_batches'] = sub
        scratch += [b1, b2]
    scratch.append(bam_path)
    ms_resident, ms_total_wall = st_res['sec_resident'] * 1000, wall * 1000
    ms_one = st_one['sec_resident'] * 1000
    n_batches = max(1, st['n_batches'])
    steps = n_batches / n_dev
    value = n_reads_timed / (ms_resident / 1000)
    e2e = n_reads_timed / (ms_total_wall / 1000)
    seed_ms = st_one['ms_stage'][2] / n_batches
    reads_per_launch = n_reads_timed / n_batches
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    # ---- seeding roofline: bytes COUNTED by the kernel in this run (every FM extension reads one occ block per rank: one
    # block when both ranks fall into the same block, else two), over the CUDA-event time of the seeding stage
    fm_ext = st_one['fm_extensions'] / n_batches
    blk = st_one['fm_block_bytes'] or 32
    bytes_layout = blk * (fm_ext + st_one['fm_two_block'] / n_batches)            # minimal sectors of the layout the kernel reads
    bytes_ref = 64 * (fm_ext + st_one['fm_two_block_ref'] / n_batches)           # the same extensions over the reference's 64-byte blocks
    achieved = bytes_layout / (seed_ms / 1000) / 1e9
    rnd = None
    traffic = None
    try:   # one ncu --set full capture of the same kernel on the same workload
        t = json.load(open(os.path.join(ROOT, 'profiles', 'r02_seed_traffic.json')))
        traffic = t['dram_bytes_read'] + t['dram_bytes_write']
    except Exception:
        pass
    if n_dev == 1:
        try:
            occ_bytes = a.genome_mb * 1000000 * 4 // 2          # the occ-block array the seeding kernel walks: 4 x genome symbols, 2 symbols per byte
            ind, chase = _native.random_sector_peak(0, occ_bytes)
            ind4, chase4 = _native.random_sector_peak(0, 4 << 30)
            rnd = {'independent_loads_gbs': ind, 'dependent_chain_per_thread_gbs': chase, 'footprint_bytes': occ_bytes, 'buffer_bytes': 1 << round(__import__('math').log2(occ_bytes)),
                   'over_4GiB': {'independent_loads_gbs': ind4, 'dependent_chain_per_thread_gbs': chase4},
                   'how': 'bsb_random_sector_peak: random 32-byte sectors of a buffer the size of the occ-block array (nearest power of two), 2048 threads per SM, 256 loads per thread, best of 3'}
        except Exception as e:  # noqa
            rnd = {'error': str(e)}
    ext_ms = st_one['ms_stage'][5] / n_batches
    cells = st_one['dp_cells_extend'] / n_batches
    sm_mhz = float(peaks.get('sm_max_mhz', 1965.0))
    lane_ops = 148 * 128 * sm_mhz * 1e6                     # integer lane-operations per second of the whole chip
    dp = {'kernel': 'k_extend_lanes + k_extend_tail (banded extension, ksw_extend2 semantics)', 'cells_per_launch': cells, 'cells_per_read': cells / reads_per_launch,
          'stage_ms_per_launch': ext_ms, 'gcups': cells / (ext_ms / 1000) / 1e9 if ext_ms else None,
          'int_lane_ops_per_s': lane_ops, 'ops_per_cell_floor': 20,
          'gcups_peak_at_floor': lane_ops / 20 / 1e9, 'frac_of_int_peak': (cells / (ext_ms / 1000)) / (lane_ops / 20) if ext_ms else None,
          'note': 'cells counted by the kernel in this run; floor = the 20 integer operations of one cell of the affine-gap recurrence with band/maximum tracking '
                  '(DESIGN.md); ncu thread-instructions per cell of the shipped kernel: profiles/r02_ncu_k_extend_lanes.md'}
    cpu = None
    parity = None
    if os.path.exists(bwa) and n_dev == 1:
        sp = min(a.cpu_sample_pairs, a.batch_pairs)
        s1 = os.path.join(work, 'cpu_s1.fq'); s2 = os.path.join(work, 'cpu_s2.fq')
        ref_sam = os.path.join(work, 'cpu_ref.sam'); my_sam = os.path.join(work, 'cpu_mine.sam')
        scratch += [s1, s2, ref_sam, my_sam]
        for src, dst in ((f1, s1), (f2, s2)):
            head_records(src, dst, sp)
        try:
            r, sec = reference_cpu_run(bwa, extra_args + ['-K', str(K_bases), db, s1, s2], 2 * sp, cores, sam_out=ref_sam)
            cpu = {'value': r, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference',
                   'sample': f'first {sp} pairs of the timed reads, oracle/_ref/bwa mem -t {cores}, {sec:.1f} s, clock from first batch read to exit'}
            # the same sample through the product (untimed): record-level identity with the reference's SAM
            fd = os.open(my_sam, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            rc, _ = mem(argv_common + [db, s1, s2], fd)
            os.close(fd)
            if rc == 0:
                parity = compare_sam(ref_sam, my_sam)
        except Exception as e:  # noqa
            cpu = cpu or {'value': None, 'unit': 'reads/s', 'cores': cores, 'kind': 'reference', 'sample': f'failed: {e}'}
    ranks.fence()
    ranks.close()
    line = {'metric': 'paired-end 150bp WGBS reads aligned/sec', 'value': value, 'unit': 'reads/s', 'n_gpus': n_dev, 'steps': steps,
            'warmup': W, 'ms_per_step': ms_resident / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32', 'data': 'synthetic', 'config': config, 'clocks': clocks,
            'e2e': {'value': e2e, 'unit': 'reads/s', 'h2d_bytes_per_step': int(st['h2d_bytes'] / steps), 'd2h_bytes_per_step': int(st['d2h_bytes'] / steps),
                    'api': ('bsb_mem_main_multi' if n_dev > 1 else 'bsb_mem_main') + ' (one FASTQ pair on the host -> one SAM stream to /dev/null)', 'wall_s': wall},
            'gpu_launches': st['kernel_launches'], 'batches_in_flight_per_gpu': int(os.environ.get('BSB_GPU_SLOTS', '3')),
            'multi_gpu': {'processes': 1, 'readers': 1, 'devices': n_dev, 'batches': n_batches, 'assignment': 'batch b -> device b mod N', 'collectives_on_data_path': 0,
                          'identity': multi_identity} if n_dev > 1 else None,
            'value_one_batch_in_flight': n_reads_timed / (ms_one / 1000),
            'roofline': {'bound': 'hbm', 'kernel': 'k_seed3 (+ k_pack4, k_seed3_finish: SMEM seeding stage)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)', 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
                         'algorithmic_bytes_per_launch': bytes_layout, 'algorithmic_bytes_per_read': bytes_layout / reads_per_launch,
                         'bytes_counted': f'in this run: {blk}-byte occ blocks x (FM extensions + extensions whose two ranks lie in different blocks)',
                         'fm_extensions_per_launch': fm_ext, 'fm_extensions_per_read': fm_ext / reads_per_launch,
                         'reference_layout': {'bytes_per_launch': bytes_ref, 'bytes_per_read': bytes_ref / reads_per_launch, 'gbs': bytes_ref / (seed_ms / 1000) / 1e9,
                                              'frac_of_copy_peak': bytes_ref / (seed_ms / 1000) / 1e9 / peak, 'survey_figure_bytes_per_read': SEED_BYTES_PER_READ},
                         'random_sector_peak': rnd,
                         'frac_of_random_sector_chain_peak': (achieved / rnd['dependent_chain_per_thread_gbs']) if rnd and rnd.get('dependent_chain_per_thread_gbs') else None,
                         'kernel_ms_per_launch': seed_ms, 'reads_per_launch': reads_per_launch},
            'roofline_dp': dp,
            'cpu_baseline': cpu, 'parity_vs_reference': parity, 'e2e_bam': bam_info, 'host_cores': cores, 'index': index_note,
            'stage_ms_per_batch': {k: v / n_batches for k, v in zip(('h2d', 'convert', 'seed', 'scan_sa', 'chain', 'extend', 'pestat', 'final'), st_one['ms_stage'])},
            'final_split_ms_per_batch': {'select': st_one['ms_select'] / n_batches, 'tasks': st_one['ms_tasks'] / n_batches, 'n_tasks': st_one['n_tasks'] // n_batches},
            'stage_note': 'CUDA-event stage times of the run with ONE batch in flight per GPU (with several in flight the stages of different batches overlap)',
            'host_busy_s': {'read': st['sec_read'], 'read_plan': st['sec_plan'], 'read_fill': st['sec_fill'], 'format': st['sec_format'], 'write': st['sec_write'], 'gpu_threads': st['sec_align']},
            'rescue': {'pairs_per_batch': st_one['rescue_pairs'] / n_batches, 'sw_jobs_per_batch': st_one['rescue_jobs'] / n_batches},
            'setup_s': {'simulate': t_sim, 'index_build': t_index}, 'index_hbm_bytes_per_gpu': idx.hbm_bytes}
    print(json.dumps(line))
    os.close(null)
    if not a.keep:
        for f in scratch:
            if os.path.exists(f):
                os.remove(f)
    return 0


if __name__ == '__main__':
    sys.exit(main())
