// elector_server -- the process behind the drop-in executables (service.h): one CUDA context per (device, matrix), created once, and
// the main functions of `poa` (poa_main.c) and `masterSplitter` (splitter_main.cpp) run on request with the client's working
// directory, argv and ELECTOR_DEVICE; their stdout / stderr bytes and exit status go back over the socket.
//   elector_server SOCKET_PATH        (started by the first client; leaves after ELECTOR_SERVICE_IDLE seconds without a request)
#include <map>
#include <string>
#include <vector>

#include <sys/select.h>

#include "../../include/elector_poa.h"
#include "service.h"

namespace {
std::map<std::string, elector_ctx *> g_ctx;   // one context per device and matrix file, kept until the server leaves
}

extern "C" int svc_cached_init(int device, const char *matrix, elector_ctx **ctx) {
  std::string key = std::to_string(device) + "|";
  if (matrix) {
    struct stat st;
    if (stat(matrix, &st) == 0) {
      char real[4096];
      key += std::string(realpath(matrix, real) ? real : matrix) + "|" + std::to_string((long long)st.st_mtime) + "|" + std::to_string((long long)st.st_size);
    } else key += std::string("?") + matrix;
  }
  const auto it = g_ctx.find(key);
  if (it != g_ctx.end()) { *ctx = it->second; return ELECTOR_OK; }
  const int rc = elector_poa_init(device, matrix, ctx);
  if (rc == ELECTOR_OK) g_ctx[key] = *ctx;
  return rc;
}
extern "C" void svc_cached_free(elector_ctx *) {}

#define ELECTOR_SERVER 1
#define elector_poa_init svc_cached_init
#define elector_poa_free svc_cached_free
#define main poa_cli_main
extern "C" {
#include "poa_main.c"
}
#undef main
#define main splitter_cli_main
#include "splitter_main.cpp"
#undef main
#undef elector_poa_init
#undef elector_poa_free

namespace {

std::string slurp(FILE *f) {
  std::string s;
  char buf[1 << 16];
  size_t n;
  rewind(f);
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
  return s;
}

void serve(int fd) {
  uint32_t head[5];
  if (svc_read_all(fd, head, sizeof head) != 0 || head[0] != SVC_MAGIC || head[4] > (1u << 26)) return;
  std::vector<char> buf(head[4] + 1, 0);
  if (svc_read_all(fd, buf.data(), head[4]) != 0) return;
  std::vector<char *> argv;
  const char *cwd = buf.data();
  size_t pos = strlen(cwd) + 1;
  for (uint32_t i = 0; i < head[3] && pos < head[4]; ++i) { argv.push_back(buf.data() + pos); pos += strlen(buf.data() + pos) + 1; }
  if (argv.size() != head[3]) return;
  argv.push_back(nullptr);
  setenv("ELECTOR_DEVICE", std::to_string((int32_t)head[2]).c_str(), 1);
  int code = 127;
  std::string out, err;
  if (chdir(cwd) != 0) err = std::string("elector_server: cannot enter ") + cwd + "\n";
  else {
    // what the program writes to stdout / stderr lands in two temporary files
    fflush(stdout); fflush(stderr);
    FILE *fo = tmpfile(), *fe = tmpfile();
    const int so = dup(1), se = dup(2);
    if (fo && fe && so >= 0 && se >= 0) {
      dup2(fileno(fo), 1); dup2(fileno(fe), 2);
      code = head[1] == SVC_KIND_POA ? poa_cli_main((int)head[3], argv.data()) : head[1] == SVC_KIND_SPLITTER ? splitter_cli_main((int)head[3], argv.data()) : 127;
      fflush(stdout); fflush(stderr);
      dup2(so, 1); dup2(se, 2);
      out = slurp(fo); err = slurp(fe);
    }
    if (so >= 0) close(so);
    if (se >= 0) close(se);
    if (fo) fclose(fo);
    if (fe) fclose(fe);
  }
  const uint32_t rep[4] = {SVC_MAGIC, (uint32_t)code, (uint32_t)out.size(), (uint32_t)err.size()};
  if (svc_write_all(fd, rep, sizeof rep) == 0 && svc_write_all(fd, out.data(), out.size()) == 0) svc_write_all(fd, err.data(), err.size());
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s SOCKET_PATH   (started by the poa / masterSplitter drop-ins, see csrc/service.h)\n", argv[0]); return 2; }
  const char *path = argv[1];
  const int probe = svc_connect(path);
  if (probe >= 0) { close(probe); return 0; }     // a server answers already
  unlink(path);
  const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
  if (ls < 0) return 1;
  struct sockaddr_un a;
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  snprintf(a.sun_path, sizeof a.sun_path, "%s", path);
  const mode_t old = umask(077);
  if (bind(ls, (struct sockaddr *)&a, sizeof a) != 0 || listen(ls, 256) != 0) return 1;
  umask(old);
  const char *idle_env = getenv("ELECTOR_SERVICE_IDLE");
  const int idle = idle_env ? atoi(idle_env) : 30;
  for (;;) {
    fd_set rd;
    FD_ZERO(&rd); FD_SET(ls, &rd);
    struct timeval tv = {idle > 0 ? idle : 30, 0};
    const int r = select(ls + 1, &rd, nullptr, nullptr, &tv);
    if (r == 0) break;                               // nobody asked for `idle` seconds
    if (r < 0) { if (errno == EINTR) continue; break; }
    const int fd = accept(ls, nullptr, nullptr);
    if (fd < 0) continue;
    serve(fd);
    close(fd);
  }
  unlink(path);
  close(ls);
  for (auto &kv : g_ctx) elector_poa_free(kv.second);
  return 0;
}
