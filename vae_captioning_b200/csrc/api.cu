// C-ABI surface of libvaecap.so (declared in include/vaecap.h).
#include "ops.h"

extern "C" const char* vc_last_error(void) { return vc::last_error().c_str(); }
extern "C" int vc_abi_version(void) { return 1; }
