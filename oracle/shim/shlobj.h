/* Win32 shell / COM names used by the reference's src/config.c:977-1018 to FIND its default config
 * file.  Test infrastructure only: the oracle always passes an explicit path, so the lookups fail
 * cleanly; what matters is that load_config()/save_config() (src/config.c:815-975) build. */
#ifndef ICW_SHIM_SHLOBJ_H
#define ICW_SHIM_SHLOBJ_H
#include <stdio.h>
#include <stdarg.h>
#include <ctype.h>
#include <string.h>
#include <stdlib.h>

typedef long HRESULT;
#define SUCCEEDED(hr) ((hr) >= 0)
#define FAILED(hr)    ((hr) < 0)
#define COINIT_APARTMENTTHREADED 2
#define CSIDL_APPDATA 0x1a
#define CSIDL_FLAG_CREATE 0x8000
#define SHGFP_TYPE_CURRENT 0
#define ERROR_ALREADY_EXISTS 183
static inline HRESULT CoInitializeEx(void *r, int f) { (void)r; (void)f; return -1; }
static inline void CoUninitialize(void) {}
static inline HRESULT SHGetFolderPath(void *w, int c, void *t, int f, char *out) { (void)w; (void)c; (void)t; (void)f; out[0] = 0; return -1; }
static inline int CreateDirectory(const char *p, void *s) { (void)p; (void)s; return 0; }
static inline unsigned GetLastError(void) { return 0; }

/* character classes and stdio under their tchar.h names */
typedef int _TINT;
#define _TEOF       EOF
#define _gettc      fgetc
#define _istcntrl(c) iscntrl((unsigned char)(c))
#define _istgraph(c) isgraph((unsigned char)(c))
#define _istblank(c) isblank((unsigned char)(c))
#define _tcschr     strchr
#define _tfopen(n, m) fopen(n, "r" [0] == (m)[0] ? "r" : "w")
#define _ftprintf   fprintf

/* MSVC printf/scanf dialect -> glibc: "I64" length modifier is "ll"; MSVC's long double is double,
 * so the reference's "%Lg" into a double* (src/config.c:536) must be "%lg" here */
static inline void icw_shim_fmt(char *dst, size_t cap, const char *src)
{
    size_t n = 0;
    while (*src && n + 3 < cap) {
        if (src[0] == 'I' && src[1] == '6' && src[2] == '4') { dst[n++] = 'l'; dst[n++] = 'l'; src += 3; }
        else if (src[0] == '%' && src[1] == 'L' && src[2] == 'g') { dst[n++] = '%'; dst[n++] = 'l'; dst[n++] = 'g'; src += 3; }
        else dst[n++] = *src++;
    }
    dst[n] = 0;
}
static inline int icw_shim_sprintf(char *buf, const char *fmt, ...)
{
    char f[128]; va_list ap; int r;
    icw_shim_fmt(f, sizeof f, fmt);
    va_start(ap, fmt); r = vsprintf(buf, f, ap); va_end(ap);
    return r;
}
static inline int icw_shim_sscanf(const char *buf, const char *fmt, ...)
{
    char f[128]; va_list ap; int r;
    icw_shim_fmt(f, sizeof f, fmt);
    va_start(ap, fmt); r = vsscanf(buf, f, ap); va_end(ap);
    return r;
}
#define _stprintf   icw_shim_sprintf
#define _stscanf    icw_shim_sscanf
#endif
