// oracle/shim/esp/io/Json.h -- TEST INFRASTRUCTURE.
// Shadows the reference's esp/io/Json.h for the oracle build of PathFinder.cpp (SURVEY 8c
// step 4): the real header chain (JsonAllTypes.h -> JsonEspTypes.h -> gfx/replay/Keyframe.h ->
// assets/Asset.h -> sensor/gfx) needs Magnum's GL headers, which cannot be built here.  Only
// NavMeshSettings::readFromJSON / writeToJSON (PF.cpp:68-93) use these four functions; they
// are off the query path and the oracle never calls them.
#ifndef ESP_IO_JSON_H_
#define ESP_IO_JSON_H_
#include <rapidjson/document.h>
#include <stdexcept>
#include <string>
namespace esp {
namespace nav {
struct NavMeshSettings;
}
namespace io {
typedef rapidjson::Document JsonDocument;
typedef rapidjson::GenericValue<rapidjson::UTF8<> > JsonGenericValue;
typedef rapidjson::MemoryPoolAllocator<> JsonAllocator;
inline JsonDocument parseJsonFile(const std::string&) {
  throw std::runtime_error("oracle shim: JSON I/O is not part of the oracle build");
}
inline bool writeJsonToFile(const JsonDocument&, const std::string&, bool = false, int = 7) {
  return false;
}
inline bool fromJsonValue(const JsonGenericValue&, esp::nav::NavMeshSettings&) { return false; }
inline JsonGenericValue toJsonValue(const esp::nav::NavMeshSettings&, JsonAllocator&) {
  return JsonGenericValue(rapidjson::kObjectType);
}
}  // namespace io
}  // namespace esp
#endif
