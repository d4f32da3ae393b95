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

// ---- host -> device staging of the source rows K1 references -----------------------------------------
// The only entry point that takes a HOST pointer.  When the vertical letterbox scale is an odd integer k
// (1920x1200 -> 640x400: k = 3) cv2's bilinear weights degenerate to (2048, 0) and K1 reads one source
// row in k; copying just those rows moves 1/k of the frame bytes over PCIe.  When the frames are
// contiguous and H == n_rows * row_step the rows of the whole batch form ONE arithmetic progression, so a
// single strided 2-D DMA covers the batch; otherwise one 2-D copy per frame.
extern "C" int b200yolo_stage_rows_h2d(const uint8_t* host_frames, int B, int H, int W, int64_t pitch,
                                       int64_t batch_stride, int row0, int row_step, int n_rows,
                                       uint8_t* dev_rows, int64_t dev_pitch, int64_t dev_batch_stride,
                                       void* stream) {
  B200_REQUIRE(host_frames && dev_rows, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && n_rows > 0 && row_step > 0 && row0 >= 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE((int64_t)row0 + (int64_t)(n_rows - 1) * row_step < H, B200YOLO_ERR_SHAPE);
  const size_t width = (size_t)W * 3;
  B200_REQUIRE(pitch >= (int64_t)width && dev_pitch >= (int64_t)width, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(batch_stride >= (int64_t)H * pitch && dev_batch_stride >= (int64_t)n_rows * dev_pitch,
               B200YOLO_ERR_SHAPE);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (batch_stride == (int64_t)H * pitch && dev_batch_stride == (int64_t)n_rows * dev_pitch &&
      H == n_rows * row_step) {
    e = cudaMemcpy2DAsync(dev_rows, (size_t)dev_pitch, host_frames + (int64_t)row0 * pitch,
                          (size_t)(pitch * row_step), width, (size_t)n_rows * B, cudaMemcpyHostToDevice, s);
    return e == cudaSuccess ? B200YOLO_OK : (int)e;
  }
  for (int b = 0; b < B; ++b) {
    e = cudaMemcpy2DAsync(dev_rows + b * dev_batch_stride, (size_t)dev_pitch,
                          host_frames + b * batch_stride + (int64_t)row0 * pitch, (size_t)(pitch * row_step), width,
                          (size_t)n_rows, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
  }
  return B200YOLO_OK;
}

// Generic strided host -> device copy (one 2-D DMA): `height` runs of `width` bytes, source runs spitch bytes
// apart, destination runs dpitch apart.  Used to stage only the class channels of a host head tensor
// (each frame's (nc, A) block is one contiguous run) when the DFL channels are read zero-copy.
extern "C" int b200yolo_copy2d_h2d(void* dev_dst, int64_t dpitch, const void* host_src, int64_t spitch,
                                   int64_t width, int64_t height, void* stream) {
  B200_REQUIRE(dev_dst && host_src, B200YOLO_ERR_NULL);
  B200_REQUIRE(width > 0 && height > 0 && dpitch >= width && spitch >= width, B200YOLO_ERR_SHAPE);
  cudaError_t e = cudaMemcpy2DAsync(dev_dst, (size_t)dpitch, host_src, (size_t)spitch, (size_t)width, (size_t)height,
                                    cudaMemcpyHostToDevice, (cudaStream_t)stream);
  return e == cudaSuccess ? B200YOLO_OK : (int)e;
}
