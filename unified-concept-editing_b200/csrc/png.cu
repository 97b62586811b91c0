// PNG writer for the generated images (the reference saves every image of a CSV row with PIL: evalscripts/generate-images-sd.py:45-46,
// file name {case_number}_{num}.png).  Host-only code.  8-bit RGB, adaptive per-row filter, and a PARALLEL deflate: the image is cut
// into stripes of rows, every stripe is filtered and deflated by its own thread (raw deflate, closed on a byte boundary with
// Z_SYNC_FLUSH; the last one with Z_FINISH), and the pieces are concatenated into one zlib stream whose Adler-32 is combined from
// the stripes' checksums — the same construction pigz uses.  Any PNG reader accepts the result (tests/test_png.py decodes it with PIL).
#include "uce_common.cuh"
#include <zlib.h>
#include <cerrno>
#include <cstring>
#include <thread>
#include <vector>

namespace uce {
namespace {

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// filter one row (3 bytes per pixel) with the type that minimises the sum of absolute (signed) residuals; out[0] = filter type
void filter_row(const unsigned char* cur, const unsigned char* prev, int nbytes, unsigned char* out, std::vector<unsigned char>& scratch) {
    constexpr int bpp = 3;
    scratch.resize((size_t)nbytes * 4);
    unsigned char* cand[4] = {scratch.data(), scratch.data() + nbytes, scratch.data() + 2 * (size_t)nbytes, scratch.data() + 3 * (size_t)nbytes};
    const int types[4] = {0, 1, 2, 4};      // None, Sub, Up, Paeth
    long best = -1; int best_i = 0;
    for (int t = 0; t < 4; ++t) {
        long sum = 0;
        unsigned char* o = cand[t];
        for (int i = 0; i < nbytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int pred = 0;
            switch (types[t]) { case 1: pred = a; break; case 2: pred = b; break; case 4: pred = paeth(a, b, c); break; default: break; }
            const unsigned char v = (unsigned char)(cur[i] - pred);
            o[i] = v;
            sum += v < 128 ? v : 256 - v;
        }
        if (best < 0 || sum < best) { best = sum; best_i = t; }
    }
    out[0] = (unsigned char)types[best_i];
    memcpy(out + 1, cand[best_i], (size_t)nbytes);
}

struct Stripe { int row0 = 0, rows = 0; std::vector<unsigned char> z; uLong adler = 1; size_t raw_len = 0; int rc = 0; };

void deflate_stripe(const unsigned char* img, int W, Stripe& s, bool last, int level) {
    const int nbytes = W * 3;
    std::vector<unsigned char> filtered((size_t)s.rows * (nbytes + 1)), scratch;
    for (int r = 0; r < s.rows; ++r) {
        const int y = s.row0 + r;
        filter_row(img + (size_t)y * nbytes, y > 0 ? img + (size_t)(y - 1) * nbytes : nullptr, nbytes, filtered.data() + (size_t)r * (nbytes + 1), scratch);
    }
    s.raw_len = filtered.size();
    s.adler = adler32(adler32(0L, Z_NULL, 0), filtered.data(), (uInt)filtered.size());
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { s.rc = -1; return; }
    s.z.resize(deflateBound(&zs, (uLong)filtered.size()) + 64);
    zs.next_in = filtered.data(); zs.avail_in = (uInt)filtered.size();
    zs.next_out = s.z.data(); zs.avail_out = (uInt)s.z.size();
    const int r = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
    if ((last && r != Z_STREAM_END) || (!last && (r != Z_OK || zs.avail_in != 0))) s.rc = -2;
    s.z.resize(zs.total_out);
    deflateEnd(&zs);
}

void put_be32(std::vector<unsigned char>& v, unsigned long x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
bool write_chunk(FILE* f, const char type[4], const unsigned char* data, size_t len) {
    unsigned char hdr[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                            (unsigned char)type[0], (unsigned char)type[1], (unsigned char)type[2], (unsigned char)type[3]};
    uLong crc = crc32(0L, Z_NULL, 0);
    crc = crc32(crc, hdr + 4, 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    const unsigned char tail[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    return fwrite(hdr, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
}

}  // namespace
}  // namespace uce

using namespace uce;

extern "C" int uce_png_write_rgb8(const char* path, const unsigned char* rgb, int H, int W, int level, int threads) {
    if (!path || !rgb || H <= 0 || W <= 0 || (long)W * 3 > (1l << 30)) { set_error("uce_png_write_rgb8: bad argument"); return UCE_E_ARG; }
    if (level < 0 || level > 9) level = 6;
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    const int min_rows = 16;                                  // below that a stripe's lost dictionary costs more than the thread buys
    if (T > (H + min_rows - 1) / min_rows) T = (H + min_rows - 1) / min_rows;
    std::vector<Stripe> st(T);
    for (int t = 0; t < T; ++t) { st[t].row0 = (int)((long)H * t / T); st[t].rows = (int)((long)H * (t + 1) / T) - st[t].row0; }
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(deflate_stripe, rgb, W, std::ref(st[t]), t == T - 1, level);
    deflate_stripe(rgb, W, st[0], T == 1, level);
    for (auto& th : pool) th.join();
    std::vector<unsigned char> idat;
    size_t total = 2 + 4;
    for (auto& s : st) { if (s.rc) { set_error("deflate failed (%d)", s.rc); return UCE_E_STATE; } total += s.z.size(); }
    idat.reserve(total);
    idat.push_back(0x78);                                     // zlib header: deflate, 32 KB window; FLG makes the pair a multiple of 31
    idat.push_back(level >= 7 ? 0xDA : (level >= 6 ? 0x9C : (level >= 2 ? 0x5E : 0x01)));
    uLong adler = adler32(0L, Z_NULL, 0);
    for (auto& s : st) {
        idat.insert(idat.end(), s.z.begin(), s.z.end());
        adler = adler32_combine(adler, s.adler, (z_off_t)s.raw_len);
    }
    put_be32(idat, adler);
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot create '%s': %s", path, strerror(errno)); return UCE_E_STATE; }
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (unsigned long)W); put_be32(ihdr, (unsigned long)H);
    const unsigned char rest[5] = {8, 2, 0, 0, 0};            // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    bool ok = fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr.data(), ihdr.size());
    for (size_t off = 0; ok && off < idat.size(); off += (1u << 30)) {
        const size_t n = idat.size() - off < (1u << 30) ? idat.size() - off : (1u << 30);
        ok = write_chunk(f, "IDAT", idat.data() + off, n);
    }
    ok = ok && write_chunk(f, "IEND", nullptr, 0);
    if (fclose(f) != 0) ok = false;
    if (!ok) { set_error("write to '%s' failed: %s", path, strerror(errno)); remove(path); return UCE_E_STATE; }
    return 0;
}
