/* LD_PRELOAD helper for GPU-box debugging: prints a native backtrace on SIGSEGV/SIGABRT/SIGBUS
 * (module+offset form, resolvable with addr2line -e <lib> <offset> on the build machine).
 * Build: gcc -shared -fPIC -O1 -o gpurun_out/segv_trace.so tools/segv_trace.c */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <string.h>
#include <unistd.h>

static void handler(int sig) {
  void* frames[64];
  const char msg[] = "\n=== native backtrace (segv_trace) ===\n";
  write(2, msg, sizeof(msg) - 1);
  int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

__attribute__((constructor)) static void install(void) {
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_handler = handler;
  sa.sa_flags = SA_NODEFER | SA_RESETHAND;
  sigaction(SIGSEGV, &sa, NULL);
  sigaction(SIGBUS, &sa, NULL);
  sigaction(SIGABRT, &sa, NULL);
}
