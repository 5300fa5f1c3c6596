// placeholder until the encode pipeline lands (see encode_engine.cu history)
#include "engine.h"
namespace jxlgpu {
EncodeResult EncodeOnGpu(const EncodeRequest&) { EncodeResult r; r.status = EncStatus::EncodeError; r.message = "encoder not built yet"; return r; }
}
