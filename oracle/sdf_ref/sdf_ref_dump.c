/* sdf_ref_dump.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A small driver of our own around the REFERENCE's SDF reader (SDF/C/src of the reference tree,
 * compiled where it lies by oracle/sdf_ref/Makefile into oracle/_ref/): it opens an SDF file with
 * sdf_open / sdf_read_blocklist / sdf_read_data exactly as the reference's tools do and exports
 * what the reader understood -- one text line of metadata per block on stdout and the raw bytes of
 * every block's data into <outdir>/<block number>.bin (meshes: <n>.<dim>.bin) -- so that the tests
 * can compare the product's SDF writer (csrc/sdf_io.cu) with the arrays it was given, bit for bit,
 * THROUGH THE REFERENCE'S OWN PARSER.  No reference source is copied; this file only calls the
 * public API of include/sdf.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdf.h"

static void dump(const char *dir, int n, int sub, const void *p, size_t bytes) {
    char path[4096];
    if (sub < 0) snprintf(path, sizeof path, "%s/%d.bin", dir, n);
    else snprintf(path, sizeof path, "%s/%d.%d.bin", dir, n, sub);
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); exit(2); }
    if (bytes && fwrite(p, bytes, 1, f) != 1) { perror("fwrite"); exit(2); }
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: sdf_ref_dump <file.sdf> <outdir>\n"); return 2; }
    sdf_file_t *h = sdf_open(argv[1], 0, SDF_READ, 0);
    if (!h) { fprintf(stderr, "sdf_open failed\n"); return 1; }
    if (sdf_read_blocklist(h)) { fprintf(stderr, "sdf_read_blocklist failed\n"); return 1; }
    printf("header step=%d time=%.17g nblocks=%d version=%d.%d code=%s restart=%d\n", h->step, h->time, h->nblocks,
           h->file_version, h->file_revision, h->code_name ? h->code_name : "", (int)h->restart_flag);
    sdf_block_t *b = h->blocklist;
    for (int n = 0; n < h->nblocks && b; ++n, b = b->next) {
        h->current_block = b;
        printf("block n=%d id=%s name=%s blocktype=%d datatype=%d ndims=%d dims=%lld,%lld,%lld units=%s mesh=%s stagger=%d "
               "species=%s geometry=%d",
               n, b->id, b->name, b->blocktype, b->datatype, b->ndims, (long long)b->dims[0], (long long)b->dims[1],
               (long long)b->dims[2], b->units ? b->units : "", b->mesh_id ? b->mesh_id : "", b->stagger,
               b->material_id ? b->material_id : "", b->geometry);
        if (b->extents && (b->blocktype == SDF_BLOCKTYPE_PLAIN_MESH || b->blocktype == SDF_BLOCKTYPE_POINT_MESH)) {
            printf(" extents=");
            for (int k = 0; k < 2 * b->ndims; ++k) printf("%s%.17g", k ? "," : "", b->extents[k]);
        }
        if (b->dim_labels && (b->blocktype == SDF_BLOCKTYPE_PLAIN_MESH || b->blocktype == SDF_BLOCKTYPE_POINT_MESH)) {
            printf(" labels=");
            for (int k = 0; k < b->ndims; ++k) printf("%s%s", k ? "," : "", b->dim_labels[k]);
        }
        if (b->blocktype == SDF_BLOCKTYPE_CONSTANT && b->datatype == SDF_DATATYPE_REAL8) {
            double v;
            memcpy(&v, b->const_value, sizeof v);
            printf(" const=%.17g", v);
        }
        printf("\n");
        if (b->blocktype == SDF_BLOCKTYPE_PLAIN_MESH || b->blocktype == SDF_BLOCKTYPE_POINT_MESH) {
            if (sdf_read_data(h)) { fprintf(stderr, "sdf_read_data failed on %s\n", b->id); return 1; }
            for (int k = 0; k < b->ndims; ++k) dump(argv[2], n, k, b->grids[k], (size_t)b->dims[k] * 8);
        } else if (b->blocktype == SDF_BLOCKTYPE_PLAIN_VARIABLE || b->blocktype == SDF_BLOCKTYPE_POINT_VARIABLE) {
            if (sdf_read_data(h)) { fprintf(stderr, "sdf_read_data failed on %s\n", b->id); return 1; }
            size_t ne = 1;
            for (int k = 0; k < b->ndims; ++k) ne *= (size_t)b->dims[k];
            dump(argv[2], n, -1, b->data, ne * 8);
        }
    }
    sdf_close(h);
    return 0;
}
