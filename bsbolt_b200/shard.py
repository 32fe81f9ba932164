"""Multi-GPU sharding of one `bsbolt Align` run: batch b goes to GPU b mod G, the index is
replicated per GPU, no collective is involved, and the per-GPU SAM parts are merged on the host in input
(batch) order. One process per GPU; each process calls the aligner with BSB_SHARD_INDEX/COUNT/PARTS set."""
import os


def shard_env(index, count, parts_path):
    env = dict(os.environ)
    env.update(BSB_SHARD_INDEX=str(index), BSB_SHARD_COUNT=str(count), BSB_SHARD_PARTS=parts_path)
    return env


def merge_shards(sam_paths, parts_paths, out):
    """Concatenates the shards' SAM chunks in batch order. sam_paths[i] / parts_paths[i] belong to shard i."""
    chunks = []
    for i, pp in enumerate(parts_paths):
        for line in open(pp):
            b, off, ln = line.split('\t')
            chunks.append((int(b), i, int(off), int(ln)))
    chunks.sort()
    files = [open(p, 'rb') for p in sam_paths]
    try:
        for b, i, off, ln in chunks:
            files[i].seek(off)
            left = ln
            while left:
                buf = files[i].read(min(left, 1 << 24))
                if not buf:
                    raise IOError(f'short read in {sam_paths[i]}')
                out.write(buf)
                left -= len(buf)
    finally:
        for f in files:
            f.close()
    return len(chunks)
