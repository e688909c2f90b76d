// qs_inst_d6.cu -- instantiates one kernel variant (see qs_variants.h); compiled in parallel with its siblings.
#include "qs_variants.h"

namespace qs {
VariantInfo variant_d6() {
  VariantInfo v = make_variant<double, 6, 0, true>("d6");

  return v;
}
}  // namespace qs
