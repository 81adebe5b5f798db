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
#include <algorithm>
#include <string>
#include <thread>
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

// read-only view of a whole file (mmap; a missing or empty file is an empty view)
struct FileView {
    const char *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    bool opened = false;
    explicit FileView(const char *path)
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) return;
        opened = true;
        struct stat st;
        if (fstat(fd, &st) == 0 && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                map = m;
                data = static_cast<const char *>(m);
                size = (size_t)st.st_size;
                madvise(m, size, MADV_SEQUENTIAL);
            }
        }
        close(fd);
    }
    ~FileView()
    {
        if (map) munmap(map, size);
    }
};

// What one worker found in its slice of the text: the codes it wrote (at `out`, which starts at the slice's own offset
// in the shared output buffer -- a slice never yields more codes than it has bytes) and where records start in them.
struct SliceResult {
    size_t begin = 0, end = 0;  // byte range of the slice
    size_t ncodes = 0;
    std::vector<uint64_t> starts;  // FASTA: code position of every header line; flat file: of every SQ line
};

unsigned worker_count(size_t bytes)
{
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    size_t per_worker = 8u << 20;  // at least 8 MB of text per worker (SWB_INGEST_SLICE_BYTES: tests cut small files)
    if (const char *env = getenv("SWB_INGEST_SLICE_BYTES")) per_worker = (size_t)std::max(1L, atol(env));
    const size_t by_size = bytes / per_worker + 1;
    return (unsigned)std::min<size_t>(std::min<size_t>(hw, 32), by_size);
}

// FASTA slice: starts at a header line (or is the whole header-less file)
void parse_fasta_slice(const char *d, const uint8_t *lut, uint8_t *out, SliceResult &r)
{
    uint8_t *w = out;
    size_t pos = r.begin;
    while (pos < r.end) {
        const char *nlp = static_cast<const char *>(memchr(d + pos, '\n', r.end - pos));
        const size_t nl = nlp ? (size_t)(nlp - d) : r.end;
        if (nl > pos && d[pos] == '>') {
            r.starts.push_back((uint64_t)(w - out));
        } else {
            for (size_t k = pos; k < nl; ++k) *w++ = lut[(unsigned char)d[k]];
        }
        pos = nl + 1;
    }
    r.ncodes = (size_t)(w - out);
}

// flat-file slice: starts at the beginning of a line that follows a "//" line (or at byte 0)
void parse_dat_slice(const char *d, const uint8_t *lut, uint8_t *out, SliceResult &r)
{
    uint8_t *w = out;
    bool in_seq = false;
    size_t pos = r.begin;
    while (pos < r.end) {
        const char *nlp = static_cast<const char *>(memchr(d + pos, '\n', r.end - pos));
        const size_t nl = nlp ? (size_t)(nlp - d) : r.end;
        const size_t len = nl - pos;
        if (len >= 2 && d[pos] == 'S' && d[pos + 1] == 'Q' && (len == 2 || d[pos + 2] == ' ')) {
            r.starts.push_back((uint64_t)(w - out));  // also closes an entry that lacks its "//"
            in_seq = true;
        } else if (len >= 2 && d[pos] == '/' && d[pos + 1] == '/') {
            in_seq = false;
        } else if (in_seq) {
            for (size_t k = pos; k < nl; ++k) {
                const char ch = d[k];
                if (ch != ' ' && ch != '\t' && ch != '\r') *w++ = lut[(unsigned char)ch];
            }
        }
        pos = nl + 1;
    }
    r.ncodes = (size_t)(w - out);
}

// Cuts [begin, size) into up to `want` slices whose starts satisfy `is_cut(pos)` (pos = first byte of a line).
template <class IsCut>
void cut_slices(const char *d, size_t begin, size_t size, unsigned want, IsCut is_cut, std::vector<SliceResult> &out)
{
    out.clear();
    size_t at = begin;
    for (unsigned i = 1; i <= want && at < size; ++i) {
        size_t stop = size;
        if (i < want) {
            size_t guess = begin + (size - begin) / want * i;
            if (guess <= at) continue;
            // next line start at or after `guess` that is a valid cut
            stop = size;
            size_t p = guess;
            while (p < size) {
                const char *nlp = static_cast<const char *>(memchr(d + p, '\n', size - p));
                if (!nlp) break;
                p = (size_t)(nlp - d) + 1;
                if (p < size && is_cut(p)) { stop = p; break; }
            }
        }
        SliceResult r;
        r.begin = at;
        r.end = stop;
        out.push_back(r);
        at = stop;
    }
}

template <class Fn>
void run_slices(std::vector<SliceResult> &slices, Fn fn)
{
    std::vector<std::thread> th;
    for (size_t i = 1; i < slices.size(); ++i) th.emplace_back([&, i]() { fn(slices[i]); });
    if (!slices.empty()) fn(slices[0]);
    for (size_t i = 0; i < th.size(); ++i) th[i].join();
}

// moves every slice's codes down to close the gaps (left to right: destinations never overtake sources)
size_t compact(uint8_t *buf, size_t buf_begin, std::vector<SliceResult> &slices, std::vector<size_t> &base)
{
    size_t at = 0;
    base.resize(slices.size());
    for (size_t i = 0; i < slices.size(); ++i) {
        base[i] = at;
        const size_t src = slices[i].begin - buf_begin;
        if (src != at && slices[i].ncodes) memmove(buf + at, buf + src, slices[i].ncodes);
        at += slices[i].ncodes;
    }
    return at;
}

bool make_lut(int preset, uint8_t *lut)
{
    char all[256];
    for (int i = 0; i < 256; ++i) all[i] = (char)i;
    return swb_encode(preset, all, 256, lut) == SWB_OK;
}

int hand_out(uint8_t *codes, size_t ncodes, const std::vector<uint64_t> &offsets, uint8_t **codes_out,
             uint64_t **offsets_out, uint32_t *n_out)
{
    uint64_t *o = (uint64_t *)malloc(sizeof(uint64_t) * offsets.size());
    uint8_t *c = codes ? (uint8_t *)realloc(codes, ncodes ? ncodes : 1) : (uint8_t *)malloc(1);
    if (!c) c = codes;  // a shrinking realloc that fails leaves the block valid
    if (!c || !o) {
        free(c);
        free(o);
        return SWB_ERR_NOMEM;
    }
    memcpy(o, offsets.data(), sizeof(uint64_t) * offsets.size());
    *codes_out = c;
    *offsets_out = o;
    *n_out = (uint32_t)(offsets.size() - 1);
    return SWB_OK;
}

}  // namespace

struct swb_dbfile {
    void *map;
    size_t bytes;
    FileHeader hdr;
};

// The file is mapped, cut at record boundaries into one slice per host thread (8 MB of text or more each), every slice
// is encoded by its own thread straight into the output buffer, and the slices are then closed up.
extern "C" int swb_read_fasta(const char *path, int preset, uint8_t **codes_out, uint64_t **offsets_out, uint32_t *n_out,
                              int32_t *first_id)
{
    if (!path || !codes_out || !offsets_out || !n_out) return SWB_ERR_ARG;
    uint8_t lut[256];
    if (!make_lut(preset, lut)) return SWB_ERR_ARG;
    FileView f(path);  // a missing file reads as empty: one empty record, like the reference parser
    const char *d = f.data;
    const size_t size = f.size;
    // first header line: text before it is dropped; without any the whole file is one record (first id -1)
    size_t h0 = size;
    bool seen_header = false;
    for (size_t pos = 0; pos < size;) {
        const char *nlp = static_cast<const char *>(memchr(d + pos, '\n', size - pos));
        const size_t nl = nlp ? (size_t)(nlp - d) : size;
        if (nl > pos && d[pos] == '>') { h0 = pos; seen_header = true; break; }
        pos = nl + 1;
    }
    const size_t begin = seen_header ? h0 : 0;
    uint8_t *buf = size > begin ? (uint8_t *)malloc(size - begin) : nullptr;
    if (size > begin && !buf) return SWB_ERR_NOMEM;
    std::vector<SliceResult> slices;
    // a header-less file is one record: any line start may cut it; otherwise slices start at header lines
    if (seen_header)
        cut_slices(d, begin, size, worker_count(size - begin), [&](size_t p) { return d[p] == '>'; }, slices);
    else
        cut_slices(d, begin, size, worker_count(size - begin), [&](size_t) { return true; }, slices);
    run_slices(slices, [&](SliceResult &r) { parse_fasta_slice(d, lut, buf + (r.begin - begin), r); });
    std::vector<size_t> base;
    const size_t ncodes = compact(buf, begin, slices, base);
    std::vector<uint64_t> offsets;
    if (seen_header) {
        for (size_t i = 0; i < slices.size(); ++i)
            for (size_t k = 0; k < slices[i].starts.size(); ++k) offsets.push_back(base[i] + slices[i].starts[k]);
    } else {
        offsets.push_back(0);
    }
    offsets.push_back(ncodes);
    if (first_id) *first_id = seen_header ? 0 : -1;
    return hand_out(buf, ncodes, offsets, codes_out, offsets_out, n_out);
}

extern "C" int swb_read_uniprot_dat(const char *path, int preset, uint8_t **codes_out, uint64_t **offsets_out,
                                    uint32_t *n_out)
{
    if (!path || !codes_out || !offsets_out || !n_out) return SWB_ERR_ARG;
    uint8_t lut[256];
    if (!make_lut(preset, lut)) return SWB_ERR_ARG;
    FileView f(path);
    if (!f.opened) return SWB_ERR_ARG;
    const char *d = f.data;
    const size_t size = f.size;
    uint8_t *buf = size ? (uint8_t *)malloc(size) : nullptr;
    if (size && !buf) return SWB_ERR_NOMEM;
    std::vector<SliceResult> slices;
    // cut after "//" lines: p is a line start whose previous line is "//"
    cut_slices(d, 0, size, worker_count(size),
               [&](size_t p) { return p >= 3 && d[p - 2] == '/' && d[p - 3] == '/' && (p == 3 || d[p - 4] == '\n'); }, slices);
    run_slices(slices, [&](SliceResult &r) { parse_dat_slice(d, lut, buf + r.begin, r); });
    std::vector<size_t> base;
    const size_t ncodes = compact(buf, 0, slices, base);
    // entries are back to back in the code stream: offsets = every start, then the end of the last entry
    std::vector<uint64_t> offsets;
    for (size_t i = 0; i < slices.size(); ++i)
        for (size_t k = 0; k < slices[i].starts.size(); ++k) offsets.push_back(base[i] + slices[i].starts[k]);
    if (offsets.empty()) offsets.push_back(0);
    else offsets.push_back(ncodes);
    return hand_out(buf, ncodes, offsets, codes_out, offsets_out, n_out);
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
    // validate the header without overflow before any offset is touched: the sizes must add up to the file size exactly,
    // the offsets must start at 0, never decrease and end at the residue count
    const uint64_t body = d->bytes - sizeof(FileHeader);
    const uint64_t *offs = reinterpret_cast<const uint64_t *>((const char *)map + sizeof(FileHeader));
    bool ok = memcmp(d->hdr.magic, kMagic, 8) == 0 && (uint64_t)d->hdr.n + 1 <= body / sizeof(uint64_t);
    if (ok) {
        const uint64_t offs_bytes = sizeof(uint64_t) * ((uint64_t)d->hdr.n + 1);
        ok = d->hdr.residues == body - offs_bytes && offs[0] == 0 && offs[d->hdr.n] == d->hdr.residues;
    }
    for (uint64_t i = 0; ok && i < d->hdr.n; ++i) ok = offs[i + 1] >= offs[i] && offs[i + 1] <= d->hdr.residues;
    if (!ok) {
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
