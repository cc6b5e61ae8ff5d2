/* The fixed-width names a reference `cpu` kernel starts with (cpu/codegen/cpp.rs:2054-2063 emits these `using` lines before the
 * headers; size_t is the platform's here because g++ rejects a conflicting redefinition). */
using uint8_t = unsigned char;
using uint16_t = unsigned short;
using uint32_t = unsigned int;
using uint64_t = unsigned long long;
using int8_t = signed char;
using int16_t = signed short;
using int32_t = signed int;
using int64_t = signed long long;
using size_t = unsigned long;
struct Accel;
