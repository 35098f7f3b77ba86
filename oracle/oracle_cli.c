/*
 * sharkmer_oracle — CLI over the CPU oracle (test infrastructure only).
 * Mirrors the counting-path flags of the reference CLI (src/cli.rs:165-320):
 *   -k <odd, 1..31> (default 19)   --chunks <n> (default 0)
 *   --histo-max <1..1000000> (default 10000)   -m/--max-reads <n>
 *   -s/--sample <name> (default "sample")   -o/--outdir <dir> (default "./")
 *   --paired   --validate-every <n>   input FASTQ files (plain or .gz)
 * and the orchestration of src/main.rs:112-131 (ingest -> consolidate -> stats).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skm_oracle.h"

int main(int argc, char **argv) {
    uint32_t k = 19, chunks = 0;
    uint64_t histo_max = 10000, max_reads = 0, validate_every = 0;
    const char *sample = "sample", *outdir = "./";
    int paired = 0;
    const char *inputs[64];
    int n_inputs = 0;
    char command[8192] = "";
    for (int i = 0; i < argc; i++) {
        if (i) strncat(command, " ", sizeof command - strlen(command) - 1);
        strncat(command, argv[i], sizeof command - strlen(command) - 1);
    }
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "-k") && i + 1 < argc) k = (uint32_t)strtoul(argv[++i], NULL, 10);
        else if (!strcmp(a, "--chunks") && i + 1 < argc) chunks = (uint32_t)strtoul(argv[++i], NULL, 10);
        else if (!strcmp(a, "--histo-max") && i + 1 < argc) histo_max = strtoull(argv[++i], NULL, 10);
        else if ((!strcmp(a, "-m") || !strcmp(a, "--max-reads")) && i + 1 < argc) max_reads = strtoull(argv[++i], NULL, 10);
        else if ((!strcmp(a, "-s") || !strcmp(a, "--sample")) && i + 1 < argc) sample = argv[++i];
        else if ((!strcmp(a, "-o") || !strcmp(a, "--outdir")) && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(a, "--validate-every") && i + 1 < argc) validate_every = strtoull(argv[++i], NULL, 10);
        else if (!strcmp(a, "--paired")) paired = 1;
        else if (n_inputs < 64) inputs[n_inputs++] = a;
    }
    /* src/cli.rs:662-673 */
    if (k < 1 || k >= 32) { fprintf(stderr, "k must be less than 32 (and at least 1), got %u\n", k); return 2; }
    if (k % 2 == 0) { fprintf(stderr, "k must be odd, got %u\n", k); return 2; }
    if (histo_max < 1 || histo_max > 1000000) { fprintf(stderr, "--histo-max must be between 1 and 1000000\n"); return 2; }
    if (n_inputs == 0 || (paired && n_inputs != 2)) { fprintf(stderr, "usage: sharkmer_oracle -k K [--chunks N] [-m N] [-s S] [-o DIR] reads.fastq[.gz] ...\n"); return 2; }

    char dir[4096];
    snprintf(dir, sizeof dir, "%s%s", outdir, (outdir[0] && outdir[strlen(outdir) - 1] != '/') ? "/" : "");

    orc_run *r = orc_run_new(k, chunks, histo_max);
    int rc = 0;
    if (paired) {
        if (max_reads > 0 && max_reads % 2) max_reads++; /* io.rs:483-485 */
        rc = orc_run_read_fastq_paired(r, inputs[0], inputs[1], max_reads, validate_every);
    } else {
        for (int i = 0; i < n_inputs; i++) {
            rc = orc_run_read_fastq(r, inputs[i], max_reads, validate_every);
            if (rc != 0) break;
        }
    }
    if (rc < 0) { fprintf(stderr, "Error: %s\n", orc_run_error(r)); return 1; }
    if ((rc = orc_run_finish_ingest(r)) < 0) { fprintf(stderr, "Error: %s\n", orc_run_error(r)); return 1; }
    if ((rc = orc_run_consolidate(r)) < 0) { fprintf(stderr, "Error: %s\n", orc_run_error(r)); return 1; }
    if (orc_run_write_histo(r, dir, sample) || orc_run_write_stats(r, dir, sample, command)) {
        fprintf(stderr, "Error: failed to write outputs to %s\n", dir);
        return 1;
    }
    fprintf(stderr, "reads %llu bases %llu kmers %llu unique %llu\n",
            (unsigned long long)orc_run_n_reads_read(r), (unsigned long long)orc_run_n_bases_read(r),
            (unsigned long long)orc_run_n_kmers_ingested(r),
            (unsigned long long)orc_counts_len(orc_run_table(r)));
    orc_run_free(r);
    return 0;
}
