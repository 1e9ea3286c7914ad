/* Declaration-only stand-in for <zstd_errors.h>; see shim/zstd.h. */
#ifndef ORACLE_SHIM_ZSTD_ERRORS_H
#define ORACLE_SHIM_ZSTD_ERRORS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int ZSTD_ErrorCode;
ZSTD_ErrorCode ZSTD_getErrorCode(size_t functionResult);
const char* ZSTD_getErrorString(ZSTD_ErrorCode code);
#ifdef __cplusplus
}
#endif
#endif
