// libb200yolo: version + error strings of the C ABI (include/b200yolo.h).
#include "common.cuh"

extern "C" int b200yolo_version(void) { return B200YOLO_VERSION; }

extern "C" const char* b200yolo_strerror(int code) {
  switch (code) {
    case B200YOLO_OK: return "ok";
    case B200YOLO_ERR_NULL: return "b200yolo: required pointer is NULL";
    case B200YOLO_ERR_SHAPE: return "b200yolo: invalid or inconsistent dimension";
    case B200YOLO_ERR_ALIGN: return "b200yolo: pointer or pitch misaligned";
    case B200YOLO_ERR_UNSUPPORTED: return "b200yolo: outside the implemented envelope";
    case B200YOLO_ERR_WORKSPACE: return "b200yolo: workspace too small (see b200yolo_workspace_bytes)";
    case B200YOLO_ERR_RANGE: return "b200yolo: threshold or value out of range";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "b200yolo: unknown error";
}
