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
#ifdef __cplusplus
extern "C"
#endif
int bsfm_shim_read_key_file(const char *filename, unsigned char **keys);
#endif
