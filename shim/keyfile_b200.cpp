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

int bsfm_shim_read_key_file(const char *filename, unsigned char **keys)
{
    std::vector<char> buf;
    if (!slurp_plain(filename, buf)) {
        const std::string gz = std::string(filename) + ".gz";          // keys2a.cpp:93-96
        if (!slurp_gz(gz.c_str(), buf)) {
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
    for (long i = 0; i < num; i++) {
        for (int t = 0; t < 4; t++) {                                     // row, col, scale, orientation: not needed
            p = skip_token(p, end);
            if (p == NULL) {
                printf("Invalid keypoint file format.");                  // keys2a.cpp:220
                delete[] out;
                return 0;
            }
        }
        for (int d = 0; d < 128; d++) {
            long v = 0;
            if (!read_uint(p, end, v)) {
                printf("Invalid keypoint file format.");
                delete[] out;
                return 0;
            }
            *q++ = (unsigned char) v;                                     // %hhu semantics (keys2a.cpp:234-247)
        }
    }
    *keys = out;
    return (int) num;
}
