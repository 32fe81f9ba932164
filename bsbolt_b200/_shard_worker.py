"""One process per GPU: `python -m bsbolt_b200._shard_worker <device> <index> <count> <sam> <parts> <log> -- <mem argv...>`"""
import os
import sys


def main():
    device, index, count, sam, parts, log = sys.argv[1:7]
    argv = sys.argv[8:]
    os.environ.update(BSB_SHARD_INDEX=index, BSB_SHARD_COUNT=count, BSB_SHARD_PARTS=parts)
    from bsbolt_b200 import _native
    with open(sam, 'wb') as fo, open(log, 'w') as fl:
        rc, _ = _native.mem_main(argv, device=int(device), out_fd=fo.fileno(), log_fd=fl.fileno())
    if rc:
        sys.stderr.write(_native.last_error() + '\n')
    sys.exit(rc)


if __name__ == '__main__':
    main()
