/* LD_PRELOAD shim: print a native backtrace when libstdc++ raises std::length_error (debugging aid, not product). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <execinfo.h>
#include <stdio.h>
#include <unistd.h>

void _ZSt20__throw_length_errorPKc(const char* what) {
  void* frames[64];
  int n = backtrace(frames, 64);
  fprintf(stderr, "\n[throw_trace] std::length_error(\"%s\") raised from:\n", what);
  backtrace_symbols_fd(frames, n, STDERR_FILENO);
  void (*real)(const char*) = (void (*)(const char*))dlsym(RTLD_NEXT, "_ZSt20__throw_length_errorPKc");
  real(what);
  __builtin_unreachable();
}
