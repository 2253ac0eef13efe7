/* keyfile_b200.h -- host-side key-file reader of the persistent KeyMatchFull CLI (shim/key_match_full_b200.cpp).
 * Same contract as the reference's  int ReadKeyFile(const char *filename, unsigned char **keys, keypt_t **info = NULL)
 * (src/keys2a.h:76-77, src/keys2a.cpp:87-113, parser :191-253 / gzip :255-323) for info == NULL:
 *   - opens <filename>, else <filename>.gz (zlib); prints "Could not open file: %s" and returns 0 if neither exists;
 *   - Lowe text format: "<num> <len>" then per key 4 floats (row, col, scale, orientation) and 128 integers 0..255;
 *   - *keys = new unsigned char[128 * num + 8] (caller delete[]s), returns num; the reference's messages for a bad
 *     header / descriptor length / truncated key, and 0, on malformed input.
 * Parsing is TOKEN based (whitespace-delimited numbers, any line breaks): it reads every well-formed Lowe file exactly
 * like the reference does -- tests/test_shim.py compares it with the reference's ReadKeyFile symbol -- but it does not
 * reproduce the reference's line structure (fgets of 7 lines per key + sscanf, keys2a.cpp:221-247): a file whose 128
 * descriptor values are not laid out 20 per line is accepted here and mis-read (or rejected) there.
 * The reference parses with fscanf/sscanf (the wall-clock bottleneck once matching runs on the GPU);
 * this reader slurps the file and converts digits by hand, and is safe to call from several threads.            */
#ifndef BSFM_KEYFILE_B200_H
#define BSFM_KEYFILE_B200_H
/* Binary key cache (row (f)2 of SURVEY.md section 8): when neither <filename> nor <filename>.gz exists the reader accepts
 * <filename>.bin / <filename>.bin.gz in the layout of the in-bundler readers ReadKeysFastBin / ReadKeysFastBinGzip
 * (src/keys.cpp:551-648): int32 num | num x keypt_t {x, y, scale, orient} | num x 128 bytes; bsfm_shim_write_key_bin writes it
 * (KeyMatchFull_b200_persistent does so for every text file it parsed when BSFM_WRITE_KEY_BIN=1: the next run, and
 * `bundler`, skip the text parse).  `info` (nullable) receives the keypt_t array (new float[4 * num], caller delete[]s). */
#ifdef __cplusplus
extern "C" {
#endif
int bsfm_shim_read_key_file(const char *filename, unsigned char **keys);
int bsfm_shim_read_key_file_info(const char *filename, unsigned char **keys, float **info);
int bsfm_shim_write_key_bin(const char *filename, int num, const unsigned char *keys, const float *info);
#ifdef __cplusplus
}
#endif
#endif
