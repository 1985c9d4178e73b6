/*
 * gsdr/util.h — linkage and exception-specification macros for the gsdr C ABI.
 * Drop-in for the reference header of the same name (ref: include/gsdr/util.h:19-29): same macro names, same
 * meaning, so reference-side sources that include <gsdr/util.h> keep compiling.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_UTIL_H_
#define GSDR_B200_INCLUDE_GSDR_UTIL_H_

#if defined(__cplusplus)
#define GSDR_C_LINKAGE extern "C"
#define GSDR_NO_EXCEPT noexcept
#else
#define GSDR_C_LINKAGE
#define GSDR_NO_EXCEPT
#endif

#endif /* GSDR_B200_INCLUDE_GSDR_UTIL_H_ */
