/* Compiled as plain C against include/scoary_b200.h and linked with libscoary_b200.so by
 * tests/test_abi.py: proves the boundary is a C ABI (no C++ in the header) and exercises the
 * entry points that need no GPU. */
#include <stdio.h>
#include <string.h>
#include "scoary_b200.h"

int main(void)
{
    if (sb_version() != SB_VERSION) { printf("version mismatch\n"); return 1; }

    /* native CSV packer on a tiny table: 2 data rows, isolate columns 3..5 */
    const char *csv = "\"Gene\",\"x\",\"y\",\"i1\",\"i2\",\"i3\"\n\"g1\",\"a\",\"b\",\"locus\",\"\",\"0\"\n\"g2\",\"c\",\"d\",\"-\",\"q\",\"r\"\n";
    int64_t starts[4], hdr_end = 0;
    int64_t n = sb_csv_row_starts(csv, (int64_t)strlen(csv), ',', starts, 4, &hdr_end);
    if (n != 2) { printf("row count %lld\n", (long long)n); return 2; }
    int32_t keep[3] = {3, 4, 5}, lead[1] = {0}, nf[2];
    uint64_t bits[2 * 2];
    int64_t ranges[2 * 1 * 2];
    if (sb_csv_pack_rows(csv, (int64_t)strlen(csv), ',', starts, 2, keep, 3, bits, 2, lead, 1, ranges, nf) != 0) return 3;
    if (bits[0] != 1ULL || bits[2] != 6ULL || nf[0] != 6) { printf("bits %llu %llu\n", (unsigned long long)bits[0], (unsigned long long)bits[2]); return 4; }
    if (strncmp(csv + ranges[0], "g1", 2) != 0 || strncmp(csv + ranges[2], "g2", 2) != 0) return 5;

    /* tree compiler: ((l0,l1),l2) */
    int32_t left[2] = {~0, 0}, right[2] = {~1, ~2}, order[3], units = -1;
    uint16_t ops[16];
    int k = sb_debug_compile_tree(left, right, 2, ops, 16, order, &units);
    if (k < 2 || order[0] != 0 || order[1] != 1 || order[2] != 2 || units != 0) { printf("compile %d\n", k); return 6; }

    /* no GPU in the build container: creating a context must fail loudly, never fall back */
    sb_ctx *ctx = NULL;
    int rc = sb_create(0, &ctx);
    if (rc == SB_OK) { sb_destroy(ctx); printf("gpu present: ok\n"); return 0; }
    if (ctx != NULL || strlen(sb_last_error(NULL)) == 0) return 7;
    printf("no gpu: %s\n", sb_last_error(NULL));
    return 0;
}
