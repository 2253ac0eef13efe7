/* oracle/ref_match_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" doorway into the UNMODIFIED reference matcher so python (ctypes) can call it.
 * Compiled against the reference headers in place (/root/reference/src/keys2a.h); calls
 * MatchKeys(int, unsigned char*, int, unsigned char*, double, int)  [src/keys2a.cpp:375-424]
 * and the tree overload                                            [src/keys2a.cpp:347-372].
 * max_pts_visit = 0 => exact kd-tree search (lib/ann_1.1_char/src/ANN.cpp:180-200);
 * 200 is the KeyMatchFull default (src/keys2a.h:101-107).
 */
#include <vector>
#include <string.h>
#include "keys2a.h"

extern "C" int ref_match_pair(int n1, unsigned char *k1, int n2, unsigned char *k2,
                              double ratio, int max_pts_visit, int *out_pairs, int cap)
{
    if (n1 <= 0 || n2 <= 0) return 0;
    std::vector<KeypointMatch> m = MatchKeys(n1, k1, n2, k2, ratio, max_pts_visit);
    int cnt = (int) m.size();
    for (int i = 0; i < cnt && i < cap; i++) {
        out_pairs[2 * i + 0] = m[i].m_idx1;
        out_pairs[2 * i + 1] = m[i].m_idx2;
    }
    return cnt;
}

/* tree overload: build once per database image, query many (KeyMatchFull.cpp:114,126) */
extern "C" void *ref_create_tree(int n, unsigned char *keys) { return (void *) CreateSearchTree(n, keys); }
extern "C" void ref_delete_tree(void *t) { delete (ANNkd_tree *) t; }
extern "C" int ref_match_tree(int n1, unsigned char *k1, void *tree, double ratio, int max_pts_visit,
                              int *out_pairs, int cap)
{
    std::vector<KeypointMatch> m = MatchKeys(n1, k1, (ANNkd_tree *) tree, ratio, max_pts_visit);
    int cnt = (int) m.size();
    for (int i = 0; i < cnt && i < cap; i++) {
        out_pairs[2 * i + 0] = m[i].m_idx1;
        out_pairs[2 * i + 1] = m[i].m_idx2;
    }
    return cnt;
}
