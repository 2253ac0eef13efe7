/* oracle/ref_keys_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" doorway into the UNMODIFIED in-bundler matcher (src/keys.cpp, used by `bundler --add_images`,
 * src/Bundle.cpp:3812-3820):
 *   MatchKeys(k1, k2, registered, ratio)            src/keys.cpp:717-810   annkPriSearch, annMaxPtsVisit(200): approximate
 *   MatchKeysExhaustive(k1, k2, registered, ratio)  src/keys.cpp:961-1057  annkSearch: exact -- the oracle of the GPU path
 * Both accept  sqrt((double) d0 / (double) d1) <= ratio  and, with `registered`, search only the keys of image 2 whose
 * m_extra >= 0 (returning their original indices).  Compiled against the reference headers in place. */
#include <vector>
#include "keys.h"

std::vector<KeypointMatch> MatchKeysExhaustive(const std::vector<KeypointWithDesc> &k1, const std::vector<KeypointWithDesc> &k2,
                                               bool registered, double ratio);

extern "C" int ref_keys_match(int n1, unsigned char *k1, int n2, unsigned char *k2, const int *extra2, int registered, double ratio,
                              int exhaustive, int *out_pairs, int cap)
{
    std::vector<KeypointWithDesc> a((size_t) n1), b((size_t) n2);
    for (int i = 0; i < n1; i++) a[i].m_d = k1 + (size_t) 128 * i;
    for (int i = 0; i < n2; i++) { b[i].m_d = k2 + (size_t) 128 * i; b[i].m_extra = extra2 ? extra2[i] : -1; }
    std::vector<KeypointMatch> m = exhaustive ? MatchKeysExhaustive(a, b, registered != 0, ratio) : MatchKeys(a, b, registered != 0, ratio);
    int cnt = (int) m.size();
    for (int i = 0; i < cnt && i < cap; i++) { out_pairs[2 * i] = m[i].m_idx1; out_pairs[2 * i + 1] = m[i].m_idx2; }
    return cnt;
}

/* the in-bundler key reader (src/keys.cpp:155-200: text, .gz, .bin, .bin.gz): descriptors and (x, y) of every key */
extern "C" int ref_read_key_file_with_desc(const char *filename, unsigned char *desc_out, float *xy_out, int cap)
{
    std::vector<KeypointWithDesc> k = ReadKeyFileWithDesc(filename, true);
    int n = (int) k.size();
    for (int i = 0; i < n && i < cap; i++) {
        for (int j = 0; j < 128; j++) desc_out[(size_t) 128 * i + j] = k[i].m_d[j];
        xy_out[2 * i] = k[i].m_x; xy_out[2 * i + 1] = k[i].m_y;
    }
    return n;
}
