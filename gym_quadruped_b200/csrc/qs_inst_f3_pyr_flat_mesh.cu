// qs_inst_f3_pyr_flat_mesh.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_f3_pyr_flat_mesh() {
  VariantInfo v = make_variant<float, 3, FEAT_CFG2, false>("f3_pyr_flat_mesh");

  return v;
}
}  // namespace qs
