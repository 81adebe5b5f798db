// Scoring presets and residue encoders of the product library (host, no CUDA).
//   SWB_SCORING_BLOSUM50_REF  the matrix of the reference's CUDA path: BLOSUM50 in the order
//                             ARNDCQEGHILKMFPSTWYVBJZX with the '*' row/column zeroed (SWSolver.cu:54-81),
//                             linear gap 2 (SWSolver.cu:7), unknown characters -> '*' (SWSolver.cu:91-120)
//   SWB_SCORING_IDENT3        the scheme of the reference's cpu.cpp: +3 equal chars, -3 otherwise, gap 2
//                             (cpu.cpp:6-8, 57-59)
#include <string.h>
#include "../../include/swb.h"
#include "swb_types.h"

static const char kLetters[] = "ARNDCQEGHILKMFPSTWYVBJZX";

// upper triangle incl. diagonal, row by row, of BLOSUM50 in kLetters order (24 x 24)
static const signed char kB50Upper[] = {
    /*A*/ 5, -2, -1, -2, -1, -1, -1, 0, -2, -1, -2, -1, -1, -3, -1, 1, 0, -3, -2, 0, -2, -2, -1, -1,
    /*R*/ 7, -1, -2, -4, 1, 0, -3, 0, -4, -3, 3, -2, -3, -3, -1, -1, -3, -1, -3, -1, -3, 0, -1,
    /*N*/ 7, 2, -2, 0, 0, 0, 1, -3, -4, 0, -2, -4, -2, 1, 0, -4, -2, -3, 5, -4, 0, -1,
    /*D*/ 8, -4, 0, 2, -1, -1, -4, -4, -1, -4, -5, -1, 0, -1, -5, -3, -4, 6, -4, 1, -1,
    /*C*/ 13, -3, -3, -3, -3, -2, -2, -3, -2, -2, -4, -1, -1, -5, -3, -1, -3, -2, -3, -1,
    /*Q*/ 7, 2, -2, 1, -3, -2, 2, 0, -4, -1, 0, -1, -1, -1, -3, 0, -3, 4, -1,
    /*E*/ 6, -3, 0, -4, -3, 1, -2, -3, -1, -1, -1, -3, -2, -3, 1, -3, 5, -1,
    /*G*/ 8, -2, -4, -4, -2, -3, -4, -2, 0, -2, -3, -3, -4, -1, -4, -2, -1,
    /*H*/ 10, -4, -3, 0, -1, -1, -2, -1, -2, -3, 2, -4, 0, -3, 0, -1,
    /*I*/ 5, 2, -3, 2, 0, -3, -3, -1, -3, -1, 4, -4, 4, -3, -1,
    /*L*/ 5, -3, 3, 1, -4, -3, -1, -2, -1, 1, -4, 4, -3, -1,
    /*K*/ 6, -2, -4, -1, 0, -1, -3, -2, -3, 0, -3, 1, -1,
    /*M*/ 7, 0, -3, -2, -1, -1, 0, 1, -3, 2, -1, -1,
    /*F*/ 8, -4, -3, -2, 1, 4, -1, -4, 1, -4, -1,
    /*P*/ 10, -1, -1, -4, -3, -3, -2, -3, -1, -1,
    /*S*/ 5, 2, -4, -2, -2, 0, -3, 0, -1,
    /*T*/ 5, -3, -2, 0, 0, -1, -1, -1,
    /*W*/ 15, 2, -3, -5, -2, -2, -1,
    /*Y*/ 8, -1, -3, -1, -2, -1,
    /*V*/ 5, -3, 2, -3, -1,
    /*B*/ 6, -4, 1, -1,
    /*J*/ 4, -3, -1,
    /*Z*/ 5, -1,
    /*X*/ -1,
};

extern "C" int swb_scoring_matrix(int preset, int8_t *m, int *gap)
{
    if (!m) return SWB_ERR_ARG;
    memset(m, 0, SWB_ALPHA * SWB_ALPHA);
    if (preset == SWB_SCORING_BLOSUM50_REF) {
        static_assert(sizeof(kB50Upper) == 24 * 25 / 2, "BLOSUM50 triangle size");
        const signed char *p = kB50Upper;
        for (int i = 0; i < 24; ++i)
            for (int j = i; j < 24; ++j) {
                m[i * SWB_ALPHA + j] = *p;
                m[j * SWB_ALPHA + i] = *p;
                ++p;
            }
    } else if (preset == SWB_SCORING_IDENT3) {
        for (int i = 0; i < SWB_ALPHA; ++i)
            for (int j = 0; j < SWB_ALPHA; ++j)
                m[i * SWB_ALPHA + j] = (i == SWB_PAD || j == SWB_PAD) ? 0 : (i == j ? 3 : -3);
    } else {
        return SWB_ERR_ARG;
    }
    if (gap) *gap = 2;
    return SWB_OK;
}

extern "C" int swb_encode(int preset, const char *text, size_t n, uint8_t *codes)
{
    if ((!text || !codes) && n) return SWB_ERR_ARG;
    uint8_t lut[256];
    if (preset == SWB_SCORING_BLOSUM50_REF) {
        memset(lut, SWB_STAR, sizeof lut);
        for (int i = 0; i < 24; ++i) lut[(unsigned char)kLetters[i]] = (uint8_t)i;
    } else if (preset == SWB_SCORING_IDENT3) {
        // raw-character comparison: every upper-case letter keeps a code of its own
        memset(lut, 30, sizeof lut);
        for (int i = 0; i < 24; ++i) lut[(unsigned char)kLetters[i]] = (uint8_t)i;
        lut[(unsigned char)'O'] = 25;
        lut[(unsigned char)'U'] = 26;
    } else {
        return SWB_ERR_ARG;
    }
    for (size_t i = 0; i < n; ++i) codes[i] = lut[(unsigned char)text[i]];
    return SWB_OK;
}
