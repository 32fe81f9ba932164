/* bwa_main.c -- a `bwa`-compatible executable over the C ABI: drop-in for bsbolt/External/BWA/bwa as located by
 * bsbolt/Utils/UtilityFunctions.py:86-100 (reference dispatch: main.c:35-63).
 *   bwa mem [options] <idxbase> <in1.fq> [in2.fq]      -> bsb_mem_main (SAM on stdout, log + BSStat on stderr)
 *   bwa index [-a bwtsw] [-b INT] <ref.fa>              -> bsb_index_build (GPU suffix array)
 * Device: environment variable BSB_DEVICE (default 0). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/bsbolt_b200.h"

int main(int argc, char **argv)
{
    int device = getenv("BSB_DEVICE") ? atoi(getenv("BSB_DEVICE")) : 0;
    if (argc < 2) {
        fprintf(stderr, "\nProgram: bwa (B200 build: %s)\nUsage:   bwa <command> [options]\n\nCommand: index         index sequences in the FASTA format\n         mem           BWA-MEM algorithm\n\n", bsb_version());
        return 1;
    }
    if (strcmp(argv[1], "mem") == 0) {
        int rc = bsb_mem_main(NULL, device, argc - 1, argv + 1, 1, 2, NULL);
        if (rc) fprintf(stderr, "%s\n", bsb_last_error());
        return rc;
    }
    if (strcmp(argv[1], "index") == 0) {
        const char *fa = argv[argc - 1];
        double ms = 0;
        int rc = bsb_index_build(fa, fa, device, &ms);
        if (rc) fprintf(stderr, "%s\n", bsb_last_error());
        else fprintf(stderr, "[bwa_index] GPU suffix array + BWT in %.1f ms\n", ms);
        return rc;
    }
    fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
    return 1;
}
