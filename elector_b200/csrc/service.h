/* service.h -- the persistent service behind the drop-in executables.
 *
 * elector/alignment.py starts one `masterSplitter`, up to 200 `poa` and 200 `Donatello` processes per round (:98-124).  A CUDA context
 * takes 0.3-1 s to create, far longer than the alignment of a shard (milliseconds): with one context per `poa` process the binary swap
 * is SLOWER than the CPU reference (profiles/r2r_dropin_*.json).  So the executables are thin clients: the first one starts
 * `elector_server` (next to it in bin/), which creates the context once and serves every later call over a Unix socket; the server
 * leaves after ELECTOR_SERVICE_IDLE seconds (default 30) without a request.  The command lines, output files, stdout / stderr bytes and
 * exit codes are those of running the program itself -- the server runs the same main function with the client's working directory,
 * argv and ELECTOR_DEVICE, and sends back what it wrote.  ELECTOR_SERVICE=0 (or a missing server binary, or any failure on the way)
 * runs the program in the client process as before.
 *
 * request : u32 magic, u32 kind, i32 device, u32 argc, u32 bytes, then cwd '\0' argv[0] '\0' ... argv[argc-1] '\0'
 * reply   : u32 magic, i32 exit code, u32 stdout bytes, u32 stderr bytes, then the bytes
 */
#ifndef ELECTOR_SERVICE_H
#define ELECTOR_SERVICE_H

#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/file.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#define SVC_MAGIC 0x454c4543u /* "ELEC" */
#define SVC_KIND_POA 1u
#define SVC_KIND_SPLITTER 2u

__attribute__((unused)) static void svc_socket_path(char *buf, size_t n)
{
  const char *e = getenv("ELECTOR_SERVICE_SOCKET");
  if (e && *e) snprintf(buf, n, "%s", e);
  else snprintf(buf, n, "/tmp/elector_b200_%u.sock", (unsigned)getuid());
}

__attribute__((unused)) static int svc_write_all(int fd, const void *p, size_t n)
{
  const char *c = (const char *)p;
  while (n) {
    ssize_t w = write(fd, c, n);
    if (w < 0) { if (errno == EINTR) continue; return -1; }
    c += w; n -= (size_t)w;
  }
  return 0;
}

__attribute__((unused)) static int svc_read_all(int fd, void *p, size_t n)
{
  char *c = (char *)p;
  while (n) {
    ssize_t r = read(fd, c, n);
    if (r < 0) { if (errno == EINTR) continue; return -1; }
    if (r == 0) return -1;
    c += r; n -= (size_t)r;
  }
  return 0;
}

__attribute__((unused)) static int svc_connect(const char *path)
{
  struct sockaddr_un a;
  int fd = socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0) return -1;
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  snprintf(a.sun_path, sizeof a.sun_path, "%s", path);
  if (connect(fd, (struct sockaddr *)&a, sizeof a) != 0) { close(fd); return -1; }
  return fd;
}

/* starts bin/elector_server (next to this executable) unless one answers already; returns a connected socket or -1 */
__attribute__((unused)) static int svc_connect_or_spawn(const char *sock)
{
  char exe[4096], lock[4200];
  ssize_t n;
  int fd = svc_connect(sock), lk, i;
  char *slash;
  if (fd >= 0) return fd;
  n = readlink("/proc/self/exe", exe, sizeof exe - 32);
  if (n <= 0) return -1;
  exe[n] = 0;
  slash = strrchr(exe, '/');
  if (!slash) return -1;
  strcpy(slash + 1, "elector_server");
  if (access(exe, X_OK) != 0) return -1;
  snprintf(lock, sizeof lock, "%s.lock", sock);
  lk = open(lock, O_CREAT | O_RDWR, 0600);
  if (lk < 0) return -1;
  flock(lk, LOCK_EX);                 /* one client starts the server, the others wait here and then find it */
  fd = svc_connect(sock);
  if (fd < 0) {
    pid_t pid = fork();
    if (pid == 0) {
      int dn;
      if (fork() != 0) _exit(0);      /* the server is nobody's child: no zombie, no wait */
      setsid();
      dn = open("/dev/null", O_RDWR);
      if (dn >= 0) { dup2(dn, 0); dup2(dn, 1); dup2(dn, 2); }
      for (i = 3; i < 256; i++) close(i);
      execl(exe, exe, sock, (char *)NULL);
      _exit(127);
    }
    if (pid > 0) {
      struct timespec ts = {0, 20 * 1000 * 1000};
      for (i = 0; i < 1500 && fd < 0; i++) { nanosleep(&ts, NULL); fd = svc_connect(sock); }   /* context creation: up to 30 s */
    }
  }
  flock(lk, LOCK_UN);
  close(lk);
  return fd;
}

/* Runs the call in the service.  1: done, *exit_code holds the program's exit status and its stdout / stderr have been written to
 * this process's; 0: no service (the caller runs the program itself). */
__attribute__((unused)) static int svc_try_call(unsigned kind, int argc, char **argv, int *exit_code)
{
  const char *on = getenv("ELECTOR_SERVICE"), *dev = getenv("ELECTOR_DEVICE");
  char sock[512], cwd[4096];
  uint32_t head[5], rep[4];
  size_t bytes, pos;
  char *buf, *out;
  int fd, i;
  if (on && on[0] == '0') return 0;
  if (!getcwd(cwd, sizeof cwd)) return 0;
  svc_socket_path(sock, sizeof sock);
  fd = svc_connect_or_spawn(sock);
  if (fd < 0) return 0;
  bytes = strlen(cwd) + 1;
  for (i = 0; i < argc; i++) bytes += strlen(argv[i]) + 1;
  buf = (char *)malloc(bytes);
  if (!buf) { close(fd); return 0; }
  pos = 0;
  memcpy(buf + pos, cwd, strlen(cwd) + 1); pos += strlen(cwd) + 1;
  for (i = 0; i < argc; i++) { memcpy(buf + pos, argv[i], strlen(argv[i]) + 1); pos += strlen(argv[i]) + 1; }
  head[0] = SVC_MAGIC; head[1] = kind; head[2] = (uint32_t)(dev ? atoi(dev) : 0); head[3] = (uint32_t)argc; head[4] = (uint32_t)bytes;
  if (svc_write_all(fd, head, sizeof head) != 0 || svc_write_all(fd, buf, bytes) != 0) { free(buf); close(fd); return 0; }
  free(buf);
  if (svc_read_all(fd, rep, sizeof rep) != 0 || rep[0] != SVC_MAGIC) { close(fd); return 0; }
  out = (char *)malloc((size_t)rep[2] + rep[3] + 1);
  if (!out || svc_read_all(fd, out, (size_t)rep[2] + rep[3]) != 0) { free(out); close(fd); return 0; }
  close(fd);
  fwrite(out, 1, rep[2], stdout);
  fwrite(out + rep[2], 1, rep[3], stderr);
  fflush(stdout); fflush(stderr);
  free(out);
  *exit_code = (int)(int32_t)rep[1];
  return 1;
}

#endif
