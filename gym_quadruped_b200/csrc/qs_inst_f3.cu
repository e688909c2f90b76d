// qs_inst_f3.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_f3() {
  VariantInfo v = make_variant<float, 3, 0, true>("f3");
  v.raycast = raycast_kernel<float>;
  return v;
}
}  // namespace qs
