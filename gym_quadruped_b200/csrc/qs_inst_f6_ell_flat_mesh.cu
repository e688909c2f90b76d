// qs_inst_f6_ell_flat_mesh.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_f6_ell_flat_mesh() {
  VariantInfo v = make_variant<float, 6, FEAT_ELL_FLAT_MESH6, false>("f6_ell_flat_mesh");

  return v;
}
}  // namespace qs
