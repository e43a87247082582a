// Miscellaneous C-ABI entry points: version, status descriptions, the per-thread last-error record.
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace {
struct LastError { int status; int line; char file[64]; char text[256]; };
thread_local LastError g_last = {0, 0, "", ""};
}  // namespace

extern "C" int air_version() { return 200; }

// Called by every `return AIR_ERR_*` of the kernel sources (macros in common.cuh) and by air_launch_status():
// remembers the status and the source line that produced it for the calling thread, returns the status unchanged.
extern "C" int air_internal_note_status(int status, const char* file, int line) {
  g_last.status = status;
  g_last.line = line;
  const char* base = file ? file : "";
  if (const char* s = strrchr(base, '/')) base = s + 1;
  strncpy(g_last.file, base, sizeof(g_last.file) - 1);
  g_last.file[sizeof(g_last.file) - 1] = 0;
  return status;
}

extern "C" const char* air_status_string(int status) {
  switch (status) {
    case 0: return "AIR_OK";
    case -1: return "AIR_ERR_ARG: invalid argument (null pointer, non-positive size, inconsistent geometry)";
    case -2: return "AIR_ERR_UNSUPPORTED: valid call outside what the kernels implement (alignment, channel count, size limit)";
    case -3: return "AIR_ERR_IO: file could not be opened or read";
    case -4: return "AIR_ERR_FORMAT: malformed or unsupported audio stream";
    case -5: return "AIR_ERR_CHECKSUM: CRC / MD5 mismatch in the audio stream";
    case -6: return "AIR_ERR_NOMEM: host allocation failed";
    case -7: return "AIR_ERR_DRIVER: cuTensorMapEncodeTiled could not be resolved (no CUDA driver loaded)";
    default: break;
  }
  if (status >= 10000) return "cuTensorMapEncodeTiled failed (status - 10000 = CUresult)";
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown status";
}

extern "C" const char* air_last_error_string() {
  if (g_last.status == 0) return "no error recorded on this thread";
  if (g_last.file[0])
    snprintf(g_last.text, sizeof(g_last.text), "status %d at %s:%d: %s", g_last.status, g_last.file, g_last.line,
             air_status_string(g_last.status));
  else
    snprintf(g_last.text, sizeof(g_last.text), "status %d: %s", g_last.status, air_status_string(g_last.status));
  return g_last.text;
}
