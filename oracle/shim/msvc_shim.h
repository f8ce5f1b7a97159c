/* Force-included (-include) when compiling the reference's encoder/compiler sources with g++.
 * Provides the handful of MSVC CRT names those sources use.  Test infrastructure only. */
#pragma once
#ifdef __cplusplus
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <ctime>
#include <strings.h>

#ifndef _countof
template <typename T, size_t N> constexpr size_t _countof_impl(T (&)[N]) { return N; }
#define _countof(a) _countof_impl(a)
#endif
#ifndef MAX_PATH
#define MAX_PATH 260
#endif
#define _stricmp strcasecmp
#define _strnicmp strncasecmp

static inline int fopen_s(FILE **fp, const char *name, const char *mode)
{ *fp = fopen(name, mode); return *fp ? 0 : 1; }

static inline int _vscprintf(const char *fmt, va_list va)
{ va_list c; va_copy(c, va); int n = vsnprintf(nullptr, 0, fmt, c); va_end(c); return n; }

static inline int vsprintf_s(char *buf, size_t n, const char *fmt, va_list va)
{ return vsnprintf(buf, n, fmt, va); }

template <size_t N> static inline int sprintf_s(char (&buf)[N], const char *fmt, ...)
{ va_list va; va_start(va, fmt); int r = vsnprintf(buf, N, fmt, va); va_end(va); return r; }
static inline int sprintf_s(char *buf, size_t n, const char *fmt, ...)
{ va_list va; va_start(va, fmt); int r = vsnprintf(buf, n, fmt, va); va_end(va); return r; }

template <size_t N> static inline int strcpy_s(char (&d)[N], const char *s)
{ strncpy(d, s, N); d[N - 1] = 0; return 0; }
static inline int strcpy_s(char *d, size_t n, const char *s)
{ strncpy(d, s, n); if (n) d[n - 1] = 0; return 0; }
template <size_t N> static inline int strcat_s(char (&d)[N], const char *s)
{ strncat(d, s, N - strlen(d) - 1); return 0; }
static inline int strcat_s(char *d, size_t n, const char *s)
{ strncat(d, s, n - strlen(d) - 1); return 0; }

static inline int localtime_s(struct tm *out, const time_t *t)
{ return localtime_r(t, out) ? 0 : 1; }
#endif
