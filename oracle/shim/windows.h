/* Minimal Win32 surface for compiling the in_cwave DSP sources on Linux.
 *
 * TEST INFRASTRUCTURE ONLY.  This header exists so that oracle/Makefile can
 * compile the reference's own C files, unmodified, from /root/reference/src
 * into oracle/_ref/ (see SURVEY.md section 8c).  Nothing under in_cwave_b200/
 * includes it.  Only the handful of Win32 names those files touch are given a
 * POSIX meaning here; everything GUI-related is left out on purpose.
 */
#ifndef ICW_SHIM_WINDOWS_H
#define ICW_SHIM_WINDOWS_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

/* ---- scalar types ------------------------------------------------------ */
typedef int                 BOOL;
typedef unsigned char       BYTE;
typedef uint16_t            WORD;
typedef uint32_t            DWORD;
typedef int32_t             LONG;
typedef uint32_t            ULONG;
typedef int64_t             LONGLONG;
typedef uint64_t            ULONGLONG;
typedef unsigned int        UINT;
typedef intptr_t            INT_PTR;
typedef uintptr_t           UINT_PTR;
typedef intptr_t            LONG_PTR;
typedef void               *LPVOID;
typedef void               *HANDLE;
typedef void               *HWND;
typedef void               *HINSTANCE;
typedef void               *HMODULE;
typedef char                CHAR;
typedef const char         *LPCSTR;
typedef char               *LPSTR;

#ifndef TRUE
#define TRUE  1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define WINAPI
#define CALLBACK
#define __cdecl
#define __stdcall
#define __inline            inline
#define __declspec(x)
#define MAX_PATH            260
#define WM_USER             0x0400

typedef union tagICW_LARGE_INTEGER { int64_t QuadPart; } LARGE_INTEGER;
typedef struct tagICW_GUID {
    uint32_t Data1; uint16_t Data2; uint16_t Data3; uint8_t Data4[8];
} GUID;

/* ---- wave format records (layout as on disk: pack(1)) ------------------- */
#pragma pack(push, 1)
typedef struct tagICW_WAVEFORMATEX {
    WORD  wFormatTag;
    WORD  nChannels;
    DWORD nSamplesPerSec;
    DWORD nAvgBytesPerSec;
    WORD  nBlockAlign;
    WORD  wBitsPerSample;
    WORD  cbSize;
} WAVEFORMATEX;
typedef struct tagICW_WAVEFORMATEXTENSIBLE {
    WAVEFORMATEX Format;
    union { WORD wValidBitsPerSample; WORD wSamplesPerBlock; WORD wReserved; } Samples;
    DWORD dwChannelMask;
    GUID  SubFormat;
} WAVEFORMATEXTENSIBLE;
#pragma pack(pop)

#define WAVE_FORMAT_PCM         0x0001
#define WAVE_FORMAT_IEEE_FLOAT  0x0003
#define WAVE_FORMAT_EXTENSIBLE  0xFFFE
#define STATIC_KSDATAFORMAT_SUBTYPE_PCM \
    0x00000001, 0x0000, 0x0010, {0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71}
#define STATIC_KSDATAFORMAT_SUBTYPE_IEEE_FLOAT \
    0x00000003, 0x0000, 0x0010, {0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71}

/* ---- critical sections: recursive pthread mutex ------------------------- */
typedef struct tagICW_CS { pthread_mutex_t m; } CRITICAL_SECTION;
static inline void InitializeCriticalSection(CRITICAL_SECTION *cs)
{
    pthread_mutexattr_t a;
    pthread_mutexattr_init(&a);
    pthread_mutexattr_settype(&a, PTHREAD_MUTEX_RECURSIVE);
    pthread_mutex_init(&cs->m, &a);
    pthread_mutexattr_destroy(&a);
}
static inline void DeleteCriticalSection(CRITICAL_SECTION *cs) { pthread_mutex_destroy(&cs->m); }
static inline void EnterCriticalSection(CRITICAL_SECTION *cs)  { pthread_mutex_lock(&cs->m); }
static inline void LeaveCriticalSection(CRITICAL_SECTION *cs)  { pthread_mutex_unlock(&cs->m); }

/* ---- interlocked ops ---------------------------------------------------- */
static inline LONG InterlockedCompareExchange(volatile LONG *d, LONG x, LONG c)
{ return __sync_val_compare_and_swap(d, c, x); }
static inline LONGLONG InterlockedCompareExchange64(volatile LONGLONG *d, LONGLONG x, LONGLONG c)
{ return __sync_val_compare_and_swap(d, c, x); }
static inline LONG InterlockedIncrement(volatile LONG *d)
{ return __sync_add_and_fetch(d, 1); }

/* ---- process / misc ----------------------------------------------------- */
static inline void FatalAppExit(UINT code, const char *msg)
{ (void)code; fputs(msg, stderr); abort(); }
static inline DWORD SleepEx(DWORD ms, BOOL alertable) { (void)ms; (void)alertable; return 0; }
static inline BOOL DisableThreadLibraryCalls(HANDLE h) { (void)h; return TRUE; }
#define DLL_PROCESS_ATTACH 1
#define DLL_THREAD_ATTACH  2
#define DLL_THREAD_DETACH  3
#define DLL_PROCESS_DETACH 0

/* ---- file I/O over stdio ------------------------------------------------ */
#define INVALID_HANDLE_VALUE   ((HANDLE)(intptr_t)-1)
#define GENERIC_READ           0x80000000u
#define FILE_SHARE_READ        1
#define OPEN_EXISTING          3
#define FILE_ATTRIBUTE_NORMAL  0x80
#define FILE_BEGIN             0
#define FILE_CURRENT           1
#define FILE_END               2

static inline HANDLE CreateFile(const char *name, DWORD access, DWORD share, void *sec,
                                DWORD disp, DWORD flags, HANDLE templ)
{
    (void)access; (void)share; (void)sec; (void)disp; (void)flags; (void)templ;
    FILE *f = fopen(name, "rb");
    return f ? (HANDLE)f : INVALID_HANDLE_VALUE;
}
static inline BOOL CloseHandle(HANDLE h) { return h && h != INVALID_HANDLE_VALUE ? fclose((FILE *)h) == 0 : FALSE; }
static inline BOOL ReadFile(HANDLE h, void *buf, DWORD n, DWORD *got, void *ovl)
{
    (void)ovl;
    size_t r = fread(buf, 1, n, (FILE *)h);
    if (got) *got = (DWORD)r;
    return ferror((FILE *)h) ? FALSE : TRUE;
}
static inline BOOL GetFileSizeEx(HANDLE h, LARGE_INTEGER *out)
{
    FILE *f = (FILE *)h;
    off_t cur = ftello(f);
    if (cur < 0 || fseeko(f, 0, SEEK_END)) return FALSE;
    out->QuadPart = ftello(f);
    return fseeko(f, cur, SEEK_SET) == 0;
}
static inline BOOL SetFilePointerEx(HANDLE h, LARGE_INTEGER dist, LARGE_INTEGER *newpos, DWORD how)
{
    FILE *f = (FILE *)h;
    int whence = how == FILE_BEGIN ? SEEK_SET : how == FILE_CURRENT ? SEEK_CUR : SEEK_END;
    if (fseeko(f, (off_t)dist.QuadPart, whence)) return FALSE;
    if (newpos) newpos->QuadPart = ftello(f);
    return TRUE;
}

/* ---- _fpclass over fpclassify ------------------------------------------- */
#define _FPCLASS_SNAN 0x0001
#define _FPCLASS_QNAN 0x0002
#define _FPCLASS_NINF 0x0004
#define _FPCLASS_NN   0x0008
#define _FPCLASS_ND   0x0010
#define _FPCLASS_NZ   0x0020
#define _FPCLASS_PZ   0x0040
#define _FPCLASS_PD   0x0080
#define _FPCLASS_PN   0x0100
#define _FPCLASS_PINF 0x0200
static inline int _fpclass(double v)
{
    int neg = signbit(v) != 0;
    switch (fpclassify(v)) {
    case FP_NAN:       return _FPCLASS_QNAN;
    case FP_INFINITE:  return neg ? _FPCLASS_NINF : _FPCLASS_PINF;
    case FP_SUBNORMAL: return neg ? _FPCLASS_ND : _FPCLASS_PD;
    case FP_ZERO:      return neg ? _FPCLASS_NZ : _FPCLASS_PZ;
    default:           return neg ? _FPCLASS_NN : _FPCLASS_PN;
    }
}

/* the plugin tables are opaque to the DSP path */
typedef struct In_Module  In_Module;   /* OUT.H in the reference tree defines Out_Module itself */

#endif /* ICW_SHIM_WINDOWS_H */
