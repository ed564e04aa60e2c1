// stub for the CUDA emulator build: the NCCL back end is never instantiated there (LocalComm only)
#pragma once
#include "../cuda_emu.h"
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5 } ncclDataType_t;
