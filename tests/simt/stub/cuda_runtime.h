#include "../simt_emu.h"
