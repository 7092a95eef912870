/* diagnostic preload: print a raw backtrace on SIGSEGV (addresses are resolved offline with addr2line against the
 * same oracle/_ref build).  Used once to find out why the reference's DEBUG=1 hook flavour crashed on the GPU box. */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
static void on_segv(int sig, siginfo_t* si, void* ctx) {
  void* bt[64];
  int n = backtrace(bt, 64);
  dprintf(2, "SEGV at %p\n", si->si_addr);
  backtrace_symbols_fd(bt, n, 2);
  FILE* f = fopen("/proc/self/maps", "r");
  char line[512];
  while (f && fgets(line, sizeof(line), f))
    if (strstr(line, "r-xp") && (strstr(line, "gemhook") || strstr(line, "gem-storm"))) dprintf(2, "MAP %s", line);
  _exit(139);
}
__attribute__((constructor)) static void init(void) {
  static char stack[1 << 16];
  stack_t ss = {.ss_sp = stack, .ss_size = sizeof(stack), .ss_flags = 0};
  sigaltstack(&ss, NULL);
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_sigaction = on_segv;
  sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, NULL);
}
