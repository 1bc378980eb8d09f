/* yh_dlpack.h - the slice of the DLPack ABI (dmlc/dlpack, struct layout of v0.8 and of the
 * "legacy" DLManagedTensor kept by v1.x) that libyolohot's *_dl entry points read.
 * Only layouts; no DLPack code is linked.  Guarded so that including the real dlpack.h
 * first is fine. */
#ifndef YH_DLPACK_H_
#define YH_DLPACK_H_
#include <stdint.h>

#ifndef DLPACK_DLPACK_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;

typedef struct {
    void *data;
    DLDevice device;
    int32_t ndim;
    DLDataType dtype;
    int64_t *shape;
    int64_t *strides;      /* NULL = compact row-major */
    uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void *manager_ctx;
    void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif /* DLPACK_DLPACK_H_ */
#endif /* YH_DLPACK_H_ */
