// oracle/shim: see Json.h next to this file.
#ifndef ESP_IO_JSONALLTYPES_H_
#define ESP_IO_JSONALLTYPES_H_
#include "esp/io/Json.h"
#endif
