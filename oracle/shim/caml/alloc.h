#include "caml_shim.h"
