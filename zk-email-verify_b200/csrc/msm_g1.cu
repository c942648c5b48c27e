// G1 instantiation of the MSM pipeline (separate translation unit so that G1 and G2 compile in parallel).
#define ZKE_MSM_G1
#include "msm.cu"
namespace zke { namespace dev { ZKE_DEFINE_CONSTANT_UPLOAD(upload_constants_msm_g1) } }
