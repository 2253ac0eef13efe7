/* shim/keys2a_b200.cpp -- link-time drop-in for the MatchKeys entry points of src/keys2a.cpp.
 *
 * Compiled AGAINST THE REFERENCE HEADERS (keys2a.h, ANN/ANN.h), it defines the two C++ overloads
 *   std::vector<KeypointMatch> MatchKeys(int, unsigned char*, ANNkd_tree*, double, int)   src/keys2a.h:104-107
 *   std::vector<KeypointMatch> MatchKeys(int, unsigned char*, int, unsigned char*, double, int)  :99-102
 * and forwards both to bsfm_match_pair (libbsfm_b200.so).  KeyMatchFull.cpp is unchanged: it still
 * calls CreateSearchTree (kept from the reference's keys2a.cpp/ANN: the tree is only a handle here) and
 * `delete tree` (KeyMatchFull.cpp:114,150); the database descriptors are recovered from the handle with
 * tree->nPoints() / tree->thePoints() (ANN.h:773-776; annAllocPts allocates one contiguous block, so
 * thePoints()[0] is the flat n x 128 array copied in keys2a.cpp:331-335).
 *
 * Build (in a checkout that has the reference sources):
 *     g++ -std=gnu++98 -O2 -I$REF/lib/ann_1.1_char/include -I$REF/src -I<repo>/include \
 *         -DBSFM_SHIM_MATCHKEYS -c shim/keys2a_b200.cpp -o keys2a_matchkeys_b200.o
 *   and compile the reference's keys2a.cpp with -DMatchKeys=MatchKeys_reference_cpu so that its own two
 *   definitions do not clash (command-line rename, no source edit), then link
 *     KeyMatchFull.o keys2a.o keys2a_matchkeys_b200.o -lANN_char -lz -L<repo>/bundler_sfm_b200 -lbsfm_b200
 *
 * max_pts_visit is accepted and ignored: the GPU search is exact (== max_pts_visit = 0), SURVEY.md F2.
 */
#ifdef BSFM_SHIM_MATCHKEYS
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include "keys2a.h"
#include "bsfm_b200.h"

static std::vector<KeypointMatch> match_gpu(int n1, unsigned char *k1, int n2, unsigned char *k2, double ratio)
{
    std::vector<KeypointMatch> matches;
    if (n1 <= 0 || n2 <= 0) return matches;
    std::vector<int> buf(2 * (size_t) n1);
    int cnt = bsfm_match_pair(k1, n1, k2, n2, ratio, &buf[0], n1);
    if (cnt < 0) {
        printf("[MatchKeys/b200] error %d: %s\n", cnt, bsfm_last_error());
        exit(1);
    }
    for (int i = 0; i < cnt; i++) matches.push_back(KeypointMatch(buf[2 * i], buf[2 * i + 1]));
    return matches;
}

std::vector<KeypointMatch> MatchKeys(int num_keys1, unsigned char *k1, ANNkd_tree *tree2, double ratio, int max_pts_visit)
{
    (void) max_pts_visit;
    return match_gpu(num_keys1, k1, tree2->nPoints(), tree2->thePoints()[0], ratio);
}

std::vector<KeypointMatch> MatchKeys(int num_keys1, unsigned char *k1, int num_keys2, unsigned char *k2, double ratio, int max_pts_visit)
{
    (void) max_pts_visit;
    return match_gpu(num_keys1, k1, num_keys2, k2, ratio);
}
#endif
