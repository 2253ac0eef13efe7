// keyfile_b200.cpp -- see keyfile_b200.h.  Host code of the shim, no CUDA.
#include "keyfile_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace {

bool slurp_plain(const char *path, std::vector<char> &buf)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t) sz : 0);
    size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    buf.resize(got);
    fclose(f);
    return true;
}

bool slurp_gz(const char *path, std::vector<char> &buf)
{
    gzFile g = gzopen(path, "rb");
    if (g == NULL) return false;
    buf.clear();
    std::vector<char> chunk(1 << 20);
    int n;
    while ((n = gzread(g, chunk.data(), (unsigned) chunk.size())) > 0) buf.insert(buf.end(), chunk.begin(), chunk.begin() + n);
    gzclose(g);
    return true;
}

inline const char *skip_ws(const char *p, const char *end)
{
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
    return p;
}
// one whitespace-delimited token; returns NULL at end of input
inline const char *skip_token(const char *p, const char *end)
{
    p = skip_ws(p, end);
    if (p >= end) return NULL;
    while (p < end && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
    return p;
}
// non-negative decimal integer (what a Lowe key file holds); false if the next token is not one
inline bool read_uint(const char *&p, const char *end, long &v)
{
    p = skip_ws(p, end);
    if (p >= end) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = (*p == '-'); p++; }
    if (p >= end || *p < '0' || *p > '9') return false;
    long x = 0;
    while (p < end && *p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); p++; }
    v = neg ? -x : x;
    return true;
}

}  // namespace

// binary key cache of the in-bundler reader (src/keys.cpp:551-648, ReadKeysFastBin / ReadKeysFastBinGzip):
//   int32 num | num x keypt_t {float x, y, scale, orient} | num x 128 descriptor bytes
static int parse_key_bin(const std::vector<char> &buf, unsigned char **keys, float **info)
{
    if (buf.size() < 4) { printf("Invalid keypoint file\n"); return 0; }
    int num = 0;
    memcpy(&num, buf.data(), 4);
    if (num < 0 || buf.size() < 4 + (size_t) num * (16 + 128)) { printf("Invalid keypoint file format."); return 0; }
    unsigned char *out = new unsigned char[128 * (size_t) num + 8];
    memcpy(out, buf.data() + 4 + (size_t) num * 16, (size_t) num * 128);
    if (info) {
        *info = new float[4 * (size_t) num + 1];
        memcpy(*info, buf.data() + 4, (size_t) num * 16);
    }
    *keys = out;
    return num;
}

int bsfm_shim_write_key_bin(const char *filename, int num, const unsigned char *keys, const float *info)
{
    FILE *f = fopen(filename, "wb");
    if (!f) return 0;
    std::vector<float> zeros;
    if (!info) { zeros.assign(4 * (size_t) (num > 0 ? num : 0), 0.f); info = zeros.data(); }
    bool ok = fwrite(&num, 4, 1, f) == 1;
    if (num > 0) ok = ok && fwrite(info, 16, (size_t) num, f) == (size_t) num && fwrite(keys, 128, (size_t) num, f) == (size_t) num;
    fclose(f);
    return ok ? 1 : 0;
}

int bsfm_shim_read_key_file(const char *filename, unsigned char **keys) { return bsfm_shim_read_key_file_info(filename, keys, NULL); }

int bsfm_shim_read_key_file_info(const char *filename, unsigned char **keys, float **info)
{
    std::vector<char> buf;
    if (!slurp_plain(filename, buf)) {
        const std::string gz = std::string(filename) + ".gz";          // keys2a.cpp:93-96
        if (!slurp_gz(gz.c_str(), buf)) {
            // the in-bundler reader also accepts a binary cache beside the text file (keys.cpp:166-190): <name>.bin, <name>.bin.gz
            const std::string bin = std::string(filename) + ".bin", bingz = bin + ".gz";
            if (slurp_plain(bin.c_str(), buf) || slurp_gz(bingz.c_str(), buf)) return parse_key_bin(buf, keys, info);
            printf("Could not open file: %s\n", filename);               // keys2a.cpp:99
            return 0;
        }
    }
    const char *p = buf.data(), *end = buf.data() + buf.size();
    long num = 0, len = 0;
    if (!read_uint(p, end, num) || !read_uint(p, end, len)) {
        printf("Invalid keypoint file\n");                               // keys2a.cpp:198
        return 0;
    }
    if (len != 128) {
        printf("Keypoint descriptor length invalid (should be 128).");   // keys2a.cpp:203
        return 0;
    }
    if (num < 0) num = 0;
    unsigned char *out = new unsigned char[128 * (size_t) num + 8];       // keys2a.cpp:207
    unsigned char *q = out;
    float *inf = info ? new float[4 * (size_t) num + 1] : NULL;
    for (long i = 0; i < num; i++) {
        float hdr[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < 4; t++) {                                     // y (row), x (col), scale, orientation (keys2a.cpp:214-219)
            const char *tok = skip_ws(p, end);
            p = skip_token(p, end);
            if (p == NULL) {
                printf("Invalid keypoint file format.");                  // keys2a.cpp:220
                delete[] out; delete[] inf;
                return 0;
            }
            if (inf) { char tmp[64]; size_t len = (size_t) (p - tok); if (len > 63) len = 63; memcpy(tmp, tok, len); tmp[len] = 0; hdr[t] = strtof(tmp, NULL); }
        }
        if (inf) { inf[4 * i + 0] = hdr[1]; inf[4 * i + 1] = hdr[0]; inf[4 * i + 2] = hdr[2]; inf[4 * i + 3] = hdr[3]; }   // keypt_t = {x, y, scale, orient}
        for (int d = 0; d < 128; d++) {
            long v = 0;
            if (!read_uint(p, end, v)) {
                printf("Invalid keypoint file format.");
                delete[] out; delete[] inf;
                return 0;
            }
            *q++ = (unsigned char) v;                                     // %hhu semantics (keys2a.cpp:234-247)
        }
    }
    *keys = out;
    if (info) *info = inf;
    return (int) num;
}
