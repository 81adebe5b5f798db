// bin/swb_mkdb: text database -> encoded database file for `bin/main --db <file>.swbdb` and swb_dbfile_open.
//   swb_mkdb <in.fasta> <out.swbdb>            multi-FASTA, record rules of the reference parser (FASTAParsers.h)
//   swb_mkdb --uniprot-dat <in.dat> <out.swbdb>  UniProt flat file, SQ blocks (the recipe of the reference's parse.py)
// Needs no GPU.
#include <stdio.h>
#include <string.h>

#include "swb.h"

int main(int argc, char **argv)
{
    const bool dat = argc == 4 && !strcmp(argv[1], "--uniprot-dat");
    if (!(argc == 3 || dat)) {
        fprintf(stderr, "usage: %s [--uniprot-dat] <in> <out.swbdb>\n", argv[0]);
        return 2;
    }
    const char *in = argv[dat ? 2 : 1], *out = argv[dat ? 3 : 2];
    uint8_t *codes = nullptr;
    uint64_t *offsets = nullptr;
    uint32_t n = 0;
    int32_t first_id = 0;
    const int rc = dat ? swb_read_uniprot_dat(in, SWB_SCORING_BLOSUM50_REF, &codes, &offsets, &n)
                       : swb_read_fasta(in, SWB_SCORING_BLOSUM50_REF, &codes, &offsets, &n, &first_id);
    if (rc != SWB_OK) {
        fprintf(stderr, "cannot read %s\n", in);
        return 1;
    }
    if (swb_dbfile_write(out, codes, offsets, n, first_id) != SWB_OK) {
        fprintf(stderr, "cannot write %s\n", out);
        return 1;
    }
    printf("%u sequences, %llu residues -> %s\n", n, (unsigned long long)offsets[n], out);
    swb_free(codes);
    swb_free(offsets);
    return 0;
}
