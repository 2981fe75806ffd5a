/* TEST INFRASTRUCTURE ONLY — forwards to the stub vkh.h */
#include "vkh.h"
