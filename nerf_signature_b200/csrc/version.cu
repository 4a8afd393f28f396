#include "nsig.h"
extern "C" const char* nsig_version(void) { return "nsig_b200 0.1.0 sm_100a"; }
