/* poa_main.c -- drop-in replacement for the reference `poa` executable
 * (src/poa-graph/main.c), a thin C shell over the C-ABI in include/elector_poa.h.
 *
 * Same command line as elector/alignment.py:60 builds:
 *   poa -pir OUT -preserve_seqorder -corrected_reads_fasta F3 -reference_reads_fasta F1
 *       -uncorrected_reads_fasta F2 -preserve_seqorder -threads 1 -pathMatrix blosum80.mat
 * Recognised flags are the five the reference parses (main.c:95,108-111); every other
 * token is ignored (main.c:85-113); argc<2 prints a usage text and exits -1 (main.c:40-83);
 * exit status 0 on success, 1 when the matrix or a FASTA file cannot be read or holds no
 * record (main.c:149-155,245-262).  stdout carries the reference's "0 1 2 " line per
 * window (buildup_lpo.c:545).  The GPU is chosen with ELECTOR_DEVICE (default 0).
 * By default the work is done by a persistent server process that the first call starts (csrc/service.h): alignment.py starts
 * a process per shard, and a CUDA context per process costs more than the reference's whole CPU run of the shard.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/elector_poa.h"
#ifndef ELECTOR_SERVER
#include "service.h"
#endif

int main(int argc, char **argv)
{
  const char *pir = "default_output_msa.fasta";
  const char *cor = NULL, *unc = NULL, *ref = NULL, *matrix = "./blosum80.mat";
  const char *dev_env = getenv("ELECTOR_DEVICE");
  elector_ctx *ctx = NULL;
  int i, rc;

  if (argc < 2) {
    fprintf(stderr,
            "\nUsage: %s [OPTIONS]\n"
            "Three-way partial-order alignment of (reference, corrected, uncorrected) read windows.\n\n"
            "  -reference_reads_fasta FILE    reference windows (FASTA)\n"
            "  -corrected_reads_fasta FILE    corrected windows (FASTA)\n"
            "  -uncorrected_reads_fasta FILE  uncorrected windows (FASTA)\n"
            "  -pathMatrix FILE               score matrix (default ./blosum80.mat)\n"
            "  -pir FILE                      output MSA in PIR format (default default_output_msa.fasta)\n\n",
            argv[0]);
    exit(-1);
  }
#ifndef ELECTOR_SERVER
  /* the persistent service runs this very function with its long-lived context (service.h); ELECTOR_SERVICE=0: here, as before */
  if (svc_try_call(SVC_KIND_POA, argc, argv, &rc)) return rc;
#endif
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-pir")) { pir = argv[++i]; continue; }
    if (!strcmp(argv[i], "-corrected_reads_fasta")) { cor = argv[++i]; continue; }
    if (!strcmp(argv[i], "-uncorrected_reads_fasta")) { unc = argv[++i]; continue; }
    if (!strcmp(argv[i], "-reference_reads_fasta")) { ref = argv[++i]; continue; }
    if (!strcmp(argv[i], "-pathMatrix")) { matrix = argv[++i]; continue; }
  }
  rc = elector_poa_init(dev_env ? atoi(dev_env) : 0, matrix, &ctx);
  if (rc != ELECTOR_OK) {
    fprintf(stderr, "poa: %s\n", elector_last_error(NULL));
    return 1;
  }
  if (!(cor && unc && ref)) { /* the reference does nothing without all three (main.c:241) */
    elector_poa_free(ctx);
    return 0;
  }
  rc = elector_poa_files(ctx, ref, cor, unc, pir, 1);
  if (rc != ELECTOR_OK) fprintf(stderr, "poa: %s\n", elector_last_error(ctx));
  elector_poa_free(ctx);
  return rc == ELECTOR_OK ? 0 : 1;
}
