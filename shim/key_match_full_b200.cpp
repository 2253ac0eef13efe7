// key_match_full_b200.cpp -- `KeyMatchFull <list.txt> <outfile> [window_radius]` on the persistent GPU matcher.
//
// Same command line, same key files and a `matches.init.txt` byte-identical to the reference tool RUN IN EXACT MODE
// (src/KeyMatchFull.cpp:57-160 with max_pts_visit = 0; the stock binary searches its kd-tree approximately with a
// 200-visit cap, keys2a.h:99-107, so on real data its table can differ from any exact matcher's), but the pair loop (:105-151) is ONE call sequence on the device-resident key
// database of libbsfm_b200.so (bsfm_keydb_create -> bsfm_match_run -> bsfm_match_fetch) instead of one MatchKeys
// call per pair, and the key files are parsed by shim/keyfile_b200.cpp on all host cores.
// The search is exact (== the reference with max_pts_visit = 0).  No CPU matching path: any library error ends
// the program with a message and a non-zero status.
#include <algorithm>
#include <chrono>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "bsfm_b200.h"
#include "keyfile_b200.h"

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// KeyMatchFull.cpp:25-56: one key file per line, blank lines skipped, the line is kept as written
int read_file_list(const char *list_in, std::vector<std::string> &key_files)
{
    FILE *fp = fopen(list_in, "r");
    if (fp == NULL) {
        printf("Error opening file %s for reading.\n", list_in);
        return 1;
    }
    char buf[512];
    while (fgets(buf, 512, fp)) {
        size_t n = strlen(buf);
        if (n > 0 && buf[n - 1] == '\n') buf[n - 1] = '\0';
        const char *start = buf;
        while (isspace((unsigned char) *start)) start++;
        if (strlen(start) == 0) continue;
        key_files.push_back(std::string(buf));
    }
    fclose(fp);
    if (key_files.size() == 0) {
        printf("No input files found in %s.\n", list_in);
        return 1;
    }
    return 0;
}

// appends the decimal digits of v (v >= 0) to out
inline void put_int(std::string &out, int v)
{
    char tmp[16];
    int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v > 0);
    while (n > 0) out.push_back(tmp[--n]);
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc != 3 && argc != 4) {
        printf("Usage: %s <list.txt> <outfile> [window_radius]\n", argv[0]);     // KeyMatchFull.cpp:64-67
        return EXIT_FAILURE;
    }
    const char *list_in = argv[1];
    const char *file_out = argv[2];
    const double ratio = 0.6;                                                       // KeyMatchFull.cpp:70
    int window_radius = -1;
    if (argc == 4) window_radius = atoi(argv[3]);

    double t0 = now_s();
    std::vector<std::string> key_files;
    if (read_file_list(list_in, key_files) != 0) return EXIT_FAILURE;
    FILE *f = fopen(file_out, "w");
    if (f == NULL) {
        printf("Could not open %s for writing.\n", file_out);
        return EXIT_FAILURE;
    }
    const int num_images = (int) key_files.size();
    std::vector<unsigned char *> keys(num_images, (unsigned char *) NULL);
    std::vector<int> num_keys(num_images, 0);

    // read all keys (KeyMatchFull.cpp:93-99), one worker per host core
    {
        unsigned nthreads = std::thread::hardware_concurrency();
        if (const char *e = getenv("BSFM_KEYREAD_THREADS")) nthreads = (unsigned) atoi(e);
        nthreads = std::max(1u, std::min(nthreads, (unsigned) num_images));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; t++)
            pool.emplace_back([&, t]() {
                // BSFM_WRITE_KEY_BIN=1: leave a binary cache <name>.bin (the layout src/keys.cpp:551-648 reads) beside every text file
                const bool write_bin = getenv("BSFM_WRITE_KEY_BIN") != NULL;
                for (int i = (int) t; i < num_images; i += (int) nthreads) {
                    if (!write_bin) { num_keys[i] = bsfm_shim_read_key_file(key_files[i].c_str(), &keys[i]); continue; }
                    float *info = NULL;
                    num_keys[i] = bsfm_shim_read_key_file_info(key_files[i].c_str(), &keys[i], &info);
                    if (num_keys[i] > 0 && info) {
                        FILE *probe = fopen(key_files[i].c_str(), "r");      // only when the TEXT file itself was the source
                        if (probe) { fclose(probe); bsfm_shim_write_key_bin((key_files[i] + ".bin").c_str(), num_keys[i], keys[i], info); }
                    }
                    delete[] info;
                }
            });
        for (auto &th : pool) th.join();
    }
    printf("[KeyMatchFull] Reading keys took %0.3fs\n", now_s() - t0);           // KeyMatchFull.cpp:101-102

    // one contiguous descriptor array + prefix offsets: the layout bsfm_keydb_create uploads once
    t0 = now_s();
    std::vector<int64_t> key_off(num_images + 1, 0);
    for (int i = 0; i < num_images; i++) key_off[i + 1] = key_off[i] + num_keys[i];
    std::vector<unsigned char> all((size_t) key_off[num_images] * 128 + 8);
    for (int i = 0; i < num_images; i++)
        if (num_keys[i] > 0) memcpy(all.data() + (size_t) key_off[i] * 128, keys[i], (size_t) num_keys[i] * 128);
    for (int i = 0; i < num_images; i++) delete[] keys[i];                         // KeyMatchFull.cpp:154-157

    // every visible GPU takes a contiguous share of the database images (BSFM_MATCH_GPUS overrides the count): the library
    // shards the key-database build and the pair loop and all-gathers the table over NCCL (bsfm_match_all_pairs_multi)
    int ngpus = 1;
    {
        const char *e = getenv("BSFM_MATCH_GPUS");
        const int visible = bsfm_device_count();
        ngpus = e ? atoi(e) : visible;
        if (ngpus < 1) ngpus = 1;
        if (visible > 0 && ngpus > visible) ngpus = visible;
        if (ngpus > num_images) ngpus = std::max(1, num_images);
    }
    const int64_t npairs = bsfm_match_num_pairs(num_images, window_radius);
    std::vector<int32_t> pair_counts((size_t) std::max<int64_t>(npairs, 1));
    std::vector<int32_t> matches;
    int64_t total = 0;
    for (int64_t cap = std::max<int64_t>(key_off[num_images], 1024);; cap *= 4) {      // grows on BSFM_ERR_CAPACITY
        matches.assign((size_t) cap * 2, 0);
        total = bsfm_match_all_pairs_multi(all.data(), key_off.data(), num_images, window_radius, ratio, ngpus, NULL,
                                           pair_counts.data(), npairs, matches.data(), cap);
        if (total != BSFM_ERR_CAPACITY || cap > ((int64_t) 1 << 33)) break;
    }
    if (total < 0) {
        printf("[KeyMatchFull/b200] error %lld: %s\n", (long long) total, bsfm_last_error());
        return EXIT_FAILURE;
    }
    printf("[KeyMatchFull/b200] %d GPU(s)\n", ngpus);
    printf("[KeyMatchFull] Matching took %0.3fs\n", now_s() - t0);
    fflush(stdout);

    // the writer of KeyMatchFull.cpp:105-142: pairs in (i ascending, j ascending) order, >= 16 matches only
    t0 = now_s();
    std::string out;
    out.reserve((size_t) total * 12 + (size_t) npairs * 4 + 64);
    int64_t pair = 0, pos = 0;
    for (int i = 0; i < num_images; i++) {
        int start_idx = 0;
        if (window_radius > 0) start_idx = std::max(i - window_radius, 0);
        for (int j = start_idx; j < i; j++, pair++) {
            const int c = pair_counts[(size_t) pair];
            if (c >= 16) {
                put_int(out, j); out.push_back(' '); put_int(out, i); out.push_back('\n');
                put_int(out, c); out.push_back('\n');
                for (int k = 0; k < c; k++) {
                    put_int(out, matches[(size_t) (pos + k) * 2]); out.push_back(' ');
                    put_int(out, matches[(size_t) (pos + k) * 2 + 1]); out.push_back('\n');
                }
            }
            pos += c;
        }
    }
    if (fwrite(out.data(), 1, out.size(), f) != out.size()) {
        printf("Could not write %s.\n", file_out);
        return EXIT_FAILURE;
    }
    fclose(f);
    printf("[KeyMatchFull/b200] %lld matches in %lld pairs; writing took %0.3fs\n", (long long) total, (long long) npairs, now_s() - t0);
    return EXIT_SUCCESS;
}
