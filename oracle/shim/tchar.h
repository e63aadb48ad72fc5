/* narrow-character TCHAR mapping; see windows.h in this directory. Test infrastructure only. */
#ifndef ICW_SHIM_TCHAR_H
#define ICW_SHIM_TCHAR_H
#include <string.h>
#include <strings.h>
typedef char TCHAR;
#define _T(x)       x
#define _tcscpy     strcpy
#define _tcscat     strcat
#define _tcslen     strlen
#define _tcsicmp    strcasecmp
#define _tcscmp     strcmp
#endif
