/* Declaration-only stand-in for <zstd.h>.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). The image ships the zstd
 * runtime (libzstd.so.1) but no development headers; this file declares exactly
 * the stable public-ABI symbols the reference's lib/zstd_codec.cpp:1-178 and
 * lib/tomahawk.cpp:9 use, so those sources can be compiled where they lie under
 * /root/reference and linked with -l:libzstd.so.1. Nothing here is copied from
 * the reference or from zstd; the prototypes follow zstd's documented C ABI. */
#ifndef ORACLE_SHIM_ZSTD_H
#define ORACLE_SHIM_ZSTD_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef ZSTD_CCtx ZSTD_CStream;
typedef ZSTD_DCtx ZSTD_DStream;
typedef struct ZSTD_inBuffer_s  { const void* src; size_t size; size_t pos; } ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s { void* dst;       size_t size; size_t pos; } ZSTD_outBuffer;
ZSTD_CCtx*    ZSTD_createCCtx(void);
size_t        ZSTD_freeCCtx(ZSTD_CCtx*);
ZSTD_DCtx*    ZSTD_createDCtx(void);
size_t        ZSTD_freeDCtx(ZSTD_DCtx*);
ZSTD_CStream* ZSTD_createCStream(void);
size_t        ZSTD_freeCStream(ZSTD_CStream*);
ZSTD_DStream* ZSTD_createDStream(void);
size_t        ZSTD_freeDStream(ZSTD_DStream*);
size_t ZSTD_initCStream(ZSTD_CStream*, int compressionLevel);
size_t ZSTD_initDStream(ZSTD_DStream*);
size_t ZSTD_compressStream(ZSTD_CStream*, ZSTD_outBuffer*, ZSTD_inBuffer*);
size_t ZSTD_flushStream(ZSTD_CStream*, ZSTD_outBuffer*);
size_t ZSTD_endStream(ZSTD_CStream*, ZSTD_outBuffer*);
size_t ZSTD_decompressStream(ZSTD_DStream*, ZSTD_outBuffer*, ZSTD_inBuffer*);
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char* ZSTD_getErrorName(size_t code);
const char* ZSTD_versionString(void);
#ifdef __cplusplus
}
#endif
#endif
