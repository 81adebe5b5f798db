// Host-side database ingest and the encoded on-disk database. No CUDA in this file.
//   swb_read_fasta        multi-FASTA with the record rules of the reference parser (FASTAParsers.h:73-136): a line whose
//                         first character is '>' starts a record, text before the first '>' is dropped when a '>'
//                         follows and is the only record (first id -1) when there is none, '\r' and blank lines are
//                         kept as they are -- but WITHOUT the '/' padding, which is score-neutral anyway
//   swb_read_uniprot_dat  UniProt flat file: the residues of every "SQ" block up to "//" (recipe of the reference's
//                         parse.py:24-35), file order
//   swb_dbfile_*          the encoded database as one memory-mappable file, so that a scan does not pay for parsing
//                         and encoding text again (the reference re-parses and re-packs on every run)
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>
#include <vector>

#include "../../include/swb.h"
#include "swb_types.h"

namespace {

struct FileHeader {
    char magic[8];      // "SWBDB\0\1\0"
    uint32_t n;         // sequences
    int32_t first_id;   // id of record 0 as the reference parser numbers it: 0, or -1 for a file without '>' lines
    uint64_t residues;
    uint64_t reserved;
};
static_assert(sizeof(FileHeader) == 32, "header is 32 bytes");
const char kMagic[8] = {'S', 'W', 'B', 'D', 'B', 0, 1, 0};

bool slurp(const char *path, std::string &data)
{
    data.clear();
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size > 0) {
        data.resize((size_t)size);
        const size_t got = fread(&data[0], 1, (size_t)size, f);
        data.resize(got);
    }
    fclose(f);
    return true;
}

int hand_out(const std::vector<uint8_t> &codes, const std::vector<uint64_t> &offsets, uint8_t **codes_out,
             uint64_t **offsets_out, uint32_t *n_out)
{
    const size_t n = offsets.size() - 1;
    uint8_t *c = (uint8_t *)malloc(codes.size() ? codes.size() : 1);
    uint64_t *o = (uint64_t *)malloc(sizeof(uint64_t) * offsets.size());
    if (!c || !o) {
        free(c);
        free(o);
        return SWB_ERR_NOMEM;
    }
    if (!codes.empty()) memcpy(c, codes.data(), codes.size());
    memcpy(o, offsets.data(), sizeof(uint64_t) * offsets.size());
    *codes_out = c;
    *offsets_out = o;
    *n_out = (uint32_t)n;
    return SWB_OK;
}

}  // namespace

struct swb_dbfile {
    void *map;
    size_t bytes;
    FileHeader hdr;
};

extern "C" int swb_read_fasta(const char *path, int preset, uint8_t **codes_out, uint64_t **offsets_out, uint32_t *n_out,
                              int32_t *first_id)
{
    if (!path || !codes_out || !offsets_out || !n_out) return SWB_ERR_ARG;
    uint8_t lut[256];
    {
        char all[256];
        for (int i = 0; i < 256; ++i) all[i] = (char)i;
        if (swb_encode(preset, all, 256, lut) != SWB_OK) return SWB_ERR_ARG;
    }
    std::string data;
    slurp(path, data);  // a missing file reads as empty: one empty record, like the reference parser
    std::vector<uint8_t> codes;
    std::vector<uint64_t> offsets(1, 0);
    codes.reserve(data.size());
    bool seen_header = false;
    size_t rec_start = 0;  // codes.size() at the start of the current record
    size_t pos = 0;
    const size_t size = data.size();
    while (pos < size) {
        size_t nl = data.find('\n', pos);
        if (nl == std::string::npos) nl = size;
        if (nl > pos && data[pos] == '>') {
            if (seen_header) offsets.push_back(codes.size());
            else codes.resize(rec_start);  // text before the first '>' is dropped
            seen_header = true;
            rec_start = codes.size();
        } else {
            for (size_t k = pos; k < nl; ++k) codes.push_back(lut[(unsigned char)data[k]]);
        }
        pos = nl + 1;
    }
    offsets.push_back(codes.size());
    if (first_id) *first_id = seen_header ? 0 : -1;
    return hand_out(codes, offsets, codes_out, offsets_out, n_out);
}

extern "C" int swb_read_uniprot_dat(const char *path, int preset, uint8_t **codes_out, uint64_t **offsets_out,
                                    uint32_t *n_out)
{
    if (!path || !codes_out || !offsets_out || !n_out) return SWB_ERR_ARG;
    uint8_t lut[256];
    {
        char all[256];
        for (int i = 0; i < 256; ++i) all[i] = (char)i;
        if (swb_encode(preset, all, 256, lut) != SWB_OK) return SWB_ERR_ARG;
    }
    std::string data;
    if (!slurp(path, data)) return SWB_ERR_ARG;
    std::vector<uint8_t> codes;
    std::vector<uint64_t> offsets(1, 0);
    bool in_seq = false;
    size_t pos = 0;
    const size_t size = data.size();
    while (pos < size) {
        size_t nl = data.find('\n', pos);
        if (nl == std::string::npos) nl = size;
        const size_t len = nl - pos;
        if (len >= 2 && data[pos] == 'S' && data[pos + 1] == 'Q' && (len == 2 || data[pos + 2] == ' ')) {
            if (in_seq) offsets.push_back(codes.size());
            in_seq = true;
        } else if (len >= 2 && data[pos] == '/' && data[pos + 1] == '/') {
            if (in_seq) offsets.push_back(codes.size());
            in_seq = false;
        } else if (in_seq) {
            for (size_t k = pos; k < nl; ++k) {
                const char ch = data[k];
                if (ch != ' ' && ch != '\t' && ch != '\r') codes.push_back(lut[(unsigned char)ch]);
            }
        }
        pos = nl + 1;
    }
    if (in_seq) offsets.push_back(codes.size());  // the last entry may lack its "//"
    return hand_out(codes, offsets, codes_out, offsets_out, n_out);
}

extern "C" void swb_free(void *p) { free(p); }

extern "C" int swb_dbfile_write(const char *path, const uint8_t *codes, const uint64_t *offsets, uint32_t n,
                                int32_t first_id)
{
    if (!path || !offsets || (!codes && n && offsets[n] != offsets[0])) return SWB_ERR_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return SWB_ERR_ARG;
    FileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.n = n;
    h.first_id = first_id;
    const uint64_t base = offsets[0];
    h.residues = offsets[n] - base;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    std::vector<uint64_t> rel(n + 1);
    for (uint32_t i = 0; i <= n; ++i) rel[i] = offsets[i] - base;
    ok = ok && fwrite(rel.data(), sizeof(uint64_t), rel.size(), f) == rel.size();
    if (h.residues) ok = ok && fwrite(codes + base, 1, h.residues, f) == h.residues;
    ok = (fclose(f) == 0) && ok;
    return ok ? SWB_OK : SWB_ERR_ARG;
}

extern "C" int swb_dbfile_open(const char *path, swb_dbfile **out)
{
    if (!path || !out) return SWB_ERR_ARG;
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return SWB_ERR_ARG;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(FileHeader)) {
        close(fd);
        return SWB_ERR_ARG;
    }
    void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return SWB_ERR_NOMEM;
    swb_dbfile *d = new swb_dbfile();
    d->map = map;
    d->bytes = (size_t)st.st_size;
    memcpy(&d->hdr, map, sizeof(FileHeader));
    const uint64_t need = sizeof(FileHeader) + sizeof(uint64_t) * ((uint64_t)d->hdr.n + 1) + d->hdr.residues;
    const uint64_t *offs = reinterpret_cast<const uint64_t *>((const char *)map + sizeof(FileHeader));
    if (memcmp(d->hdr.magic, kMagic, 8) != 0 || need != d->bytes || offs[0] != 0 || offs[d->hdr.n] != d->hdr.residues) {
        munmap(map, d->bytes);
        delete d;
        return SWB_ERR_ARG;
    }
    *out = d;
    return SWB_OK;
}

extern "C" uint32_t swb_dbfile_count(const swb_dbfile *d) { return d ? d->hdr.n : 0; }
extern "C" int32_t swb_dbfile_first_id(const swb_dbfile *d) { return d ? d->hdr.first_id : 0; }
extern "C" const uint64_t *swb_dbfile_offsets(const swb_dbfile *d)
{
    return d ? reinterpret_cast<const uint64_t *>((const char *)d->map + sizeof(FileHeader)) : nullptr;
}
extern "C" const uint8_t *swb_dbfile_codes(const swb_dbfile *d)
{
    return d ? reinterpret_cast<const uint8_t *>((const char *)d->map + sizeof(FileHeader) +
                                                 sizeof(uint64_t) * ((size_t)d->hdr.n + 1))
             : nullptr;
}
extern "C" void swb_dbfile_close(swb_dbfile *d)
{
    if (!d) return;
    munmap(d->map, d->bytes);
    delete d;
}
