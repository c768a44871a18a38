/* TEST INFRASTRUCTURE ONLY (oracle build shim): see sundials/sundials_types.h */
#include "sundials/sundials_types.h"
