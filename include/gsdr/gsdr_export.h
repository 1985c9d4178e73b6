/*
 * gsdr/gsdr_export.h — symbol-visibility macros.
 * The reference generates this file with CMake's generate_export_header (ref: CMakeLists.txt:52-65); it is
 * written by hand here so that <gsdr/fir.h> resolves without running a configure step.
 */
#ifndef GSDR_B200_INCLUDE_GSDR_EXPORT_H_
#define GSDR_B200_INCLUDE_GSDR_EXPORT_H_

#if defined(GSDR_STATIC_BUILD)
#define GSDR_PUBLIC
#define GSDR_PRIVATE
#elif defined(_WIN32)
#if defined(gsdr_EXPORTS)
#define GSDR_PUBLIC __declspec(dllexport)
#else
#define GSDR_PUBLIC __declspec(dllimport)
#endif
#define GSDR_PRIVATE
#else
#define GSDR_PUBLIC __attribute__((visibility("default")))
#define GSDR_PRIVATE __attribute__((visibility("hidden")))
#endif

#endif /* GSDR_B200_INCLUDE_GSDR_EXPORT_H_ */
