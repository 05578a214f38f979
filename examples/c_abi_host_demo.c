/* A plain-C consumer of libdeeprec_b200.so: what a binding in any FFI-capable host language does.
 * Uses only HOST entry points, so it runs without a GPU: library version, FarmHash Fingerprint64 / hash buckets
 * (categorical_column_with_hash_bucket), CRC-32C, and the argument validation every compute entry performs before
 * it launches anything.
 *
 *   gcc -std=c99 -I include examples/c_abi_host_demo.c -L deep_recommenders_b200/lib -ldeeprec_b200 \
 *       -Wl,-rpath,$PWD/deep_recommenders_b200/lib -o /tmp/c_abi_host_demo && /tmp/c_abi_host_demo
 */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>

#include "deeprec_b200.h"

int main(void) {
  printf("version %d\n", dr_version());
  const char* words[3] = {"Hello", "TensorFlow", "2.x"};
  uint8_t bytes[64];
  int64_t offsets[4] = {0, 0, 0, 0}, ids[3];
  int64_t n = 0;
  for (int i = 0; i < 3; ++i) {
    memcpy(bytes + n, words[i], strlen(words[i]));
    n += (int64_t)strlen(words[i]);
    offsets[i + 1] = n;
  }
  if (dr_hash_bucket_bytes_host(bytes, offsets, 3, 3, ids) != DR_OK) return 1;
  printf("to_hash_bucket_fast %" PRId64 " %" PRId64 " %" PRId64 "\n", ids[0], ids[1], ids[2]);
  printf("fingerprint64(abc) %" PRIu64 "\n", dr_fingerprint64_host((const uint8_t*)"abc", 3));
  printf("crc32c(123456789) %08x\n", (unsigned)dr_crc32c_host((const uint8_t*)"123456789", 9));
  int64_t values[2] = {6040, -1}, out[2];
  if (dr_hash_bucket_i64_host(values, 2, 1000, out) != DR_OK) return 1;
  printf("hash_bucket_i64 %" PRId64 " %" PRId64 "\n", out[0], out[1]);
  /* a compute entry rejects a bad call on the host, before any CUDA call: D = 6 is not a multiple of 4 */
  int rc = dr_gather_fwd((const float*)16, 10, (const void*)16, 8, 4, 6, (float*)16, NULL);
  printf("dr_gather_fwd(D=6) rc=%d msg=%s\n", rc, dr_last_error());
  return rc == DR_EINVAL ? 0 : 1;
}
