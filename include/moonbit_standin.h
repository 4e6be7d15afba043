/*
 * moonbit_standin.h — container stand-in for the three things the reference's C stub
 * uses from the real `moonbit.h` (which is NOT present in this image):
 *
 *   moonbit_bytes_t                      (reference: src/duckdb_native.c:42  return type)
 *   moonbit_make_bytes_raw(int32 len)    (reference: src/duckdb_native.c:43,2372,2586 ...)
 *   Moonbit_array_length(bytes)          (reference: src/duckdb_native.c:55)
 *
 * Layout used here: an 8-byte object header placed *before* the payload pointer:
 *   [int32 refcount][uint32 length] payload...
 * A real MoonBit toolchain (pinned 0.1.20260409 in the reference's .tool-versions) ships its
 * own header; when building inside a MoonBit project define DMB_HAVE_REAL_MOONBIT_H and this
 * file forwards to <moonbit.h> so the shim links against the real runtime allocator.
 * ABI compatibility with the real header layout is NOT claimed from this container.
 */
#ifndef DMB_MOONBIT_STANDIN_H
#define DMB_MOONBIT_STANDIN_H

#ifdef DMB_HAVE_REAL_MOONBIT_H
#include <moonbit.h>
#else

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t *moonbit_bytes_t;

struct dmb_moonbit_object_header {
  int32_t rc;
  uint32_t len;
};

static inline moonbit_bytes_t moonbit_make_bytes_raw(int32_t len) {
  if (len < 0) len = 0;
  struct dmb_moonbit_object_header *h = (struct dmb_moonbit_object_header *)malloc(
      sizeof(struct dmb_moonbit_object_header) + (size_t)len + 1);
  if (!h) return NULL;
  h->rc = 1;
  h->len = (uint32_t)len;
  return (moonbit_bytes_t)(h + 1);
}

#define Moonbit_array_length(obj) \
  ((int32_t)(((struct dmb_moonbit_object_header *)(obj)) - 1)->len)

/* test harness only: the real runtime frees through its RC machinery */
static inline void moonbit_standin_free(moonbit_bytes_t b) {
  if (b) free(((struct dmb_moonbit_object_header *)b) - 1);
}

#ifdef __cplusplus
}
#endif

#endif /* DMB_HAVE_REAL_MOONBIT_H */
#endif /* DMB_MOONBIT_STANDIN_H */
