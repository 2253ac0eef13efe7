// shim/keys_b200.cpp -- link-time replacement of the in-bundler matcher of src/keys.cpp (row (f)4 of SURVEY.md section 8):
//   std::vector<KeypointMatch> MatchKeys(const std::vector<KeypointWithDesc> &k1, const std::vector<KeypointWithDesc> &k2,
//                                        bool registered = false, double ratio = 0.6);              src/keys.h, keys.cpp:717-810
//   std::vector<KeypointMatch> MatchKeysExhaustive(...same...);                                      keys.cpp:961-1057
// called by BundlerApp::BundleRegisterImage for `--add_images` (src/Bundle.cpp:3812-3820).  Compiled against the REFERENCE
// headers (shim/Makefile).  Both forward to bsfm_match_pair_test(..., BSFM_RATIO_TEST_KEYS, ...): exact 2-NN on the GPU with
// the acceptance test sqrt(d0 / d1) <= ratio of keys.cpp:786; `registered` restricts image 2 to keys with m_extra >= 0 and
// maps the indices back, like keys.cpp:727-737, :790-796.  The stock MatchKeys searches approximately (200-visit cap): this
// one is exact for both names and says so.  Library errors end the program like the reference's fatal paths (no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "keys.h"
#include "bsfm_b200.h"

static std::vector<KeypointMatch> match_keys_b200(const std::vector<KeypointWithDesc> &k1, const std::vector<KeypointWithDesc> &k2,
                                                  bool registered, double ratio)
{
    std::vector<KeypointMatch> matches;
    std::vector<int> idx2;                                   // database row -> index in k2
    idx2.reserve(k2.size());
    for (int i = 0; i < (int) k2.size(); i++)
        if (!registered || k2[i].m_extra >= 0) idx2.push_back(i);
    const int n1 = (int) k1.size(), n2 = (int) idx2.size();
    if (n1 > 0 && n2 > 0) {
        // descriptors live behind one pointer per key: gather them into the contiguous arrays the C ABI takes
        std::vector<unsigned char> a((size_t) n1 * 128), b((size_t) n2 * 128);
        for (int i = 0; i < n1; i++) memcpy(&a[(size_t) i * 128], k1[i].m_d, 128);
        for (int i = 0; i < n2; i++) memcpy(&b[(size_t) i * 128], k2[idx2[i]].m_d, 128);
        std::vector<int32_t> out((size_t) n1 * 2);
        const int n = bsfm_match_pair_test(a.data(), n1, b.data(), n2, ratio, BSFM_RATIO_TEST_KEYS, out.data(), n1);
        if (n < 0) {
            printf("[MatchKeys/b200] error %d: %s\n", n, bsfm_last_error());
            exit(1);
        }
        matches.reserve((size_t) n);
        for (int q = 0; q < n; q++) matches.push_back(KeypointMatch(out[(size_t) 2 * q], idx2[out[(size_t) 2 * q + 1]]));
    }
    printf("[MatchKeys] Found %d matches\n", (int) matches.size());        // keys.cpp:801
    return matches;
}

std::vector<KeypointMatch> MatchKeys(const std::vector<KeypointWithDesc> &k1, const std::vector<KeypointWithDesc> &k2, bool registered, double ratio)
{
    return match_keys_b200(k1, k2, registered, ratio);
}
std::vector<KeypointMatch> MatchKeysExhaustive(const std::vector<KeypointWithDesc> &k1, const std::vector<KeypointWithDesc> &k2, bool registered, double ratio)
{
    return match_keys_b200(k1, k2, registered, ratio);
}

// doorway for the tests (ctypes cannot build std::vector<KeypointWithDesc>): same shape as oracle/ref_keys_wrap.cpp
extern "C" int shim_keys_match(int n1, unsigned char *k1, int n2, unsigned char *k2, const int *extra2, int registered, double ratio,
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
