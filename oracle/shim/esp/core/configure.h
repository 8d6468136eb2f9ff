// oracle/shim: stands in for the header cmake generates from src/esp/core/configure.h.cmake
// (every option off: no assimp, CUDA, bullet or background renderer in the oracle build).
#ifndef ESP_CORE_CONFIGURE_H_
#define ESP_CORE_CONFIGURE_H_
#endif
