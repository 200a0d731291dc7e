// forwarding header for the oracle TF shim (test infrastructure only)
#include "tf_shim.h"
