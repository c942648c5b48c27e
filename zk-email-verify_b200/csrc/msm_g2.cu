// G2 instantiation of the MSM pipeline.
#define ZKE_MSM_G2
#include "msm.cu"
namespace zke { namespace dev { ZKE_DEFINE_CONSTANT_UPLOAD(upload_constants_msm_g2) } }
