# Include paths and the PUBLIC compile definitions LuisaCompute's build gives every target that uses its headers
# (LC/src/ext/EASTL/CMakeLists.txt:25-53, LC/src/ext/CMakeLists.txt:144): a backend module must be compiled with the same ones or
# its view of the EASTL containers / allocator differs from the host program's.  Shared by csrc/Makefile (cpp_backend) and
# oracle/Makefile (cpp_host).  Development container only.
REF ?= /root/reference/luisa_compute_sys/LuisaCompute
REF_INCS := -I $(REF)/include -I $(REF)/src/ext/EASTL/include -I $(REF)/src/ext/EASTL/packages/EABase/include/Common \
            -I $(REF)/src/ext/spdlog/include -I $(REF)/src/ext/xxHash -I $(REF)/src/ext/magic_enum/include -I $(REF)/src/ext/half/include
REF_DEFS := -DEA_PRAGMA_ONCE_SUPPORTED=1 -DEA_HAVE_CPP11_CONTAINERS=1 -DEA_HAVE_CPP11_ATOMIC=1 -DEA_HAVE_CPP11_CONDITION_VARIABLE=1 \
            -DEA_HAVE_CPP11_MUTEX=1 -DEA_HAVE_CPP11_THREAD=1 -DEA_HAVE_CPP11_FUTURE=1 -DEA_HAVE_CPP11_TYPE_TRAITS=1 -DEA_HAVE_CPP11_TUPLES=1 \
            -DEA_HAVE_CPP11_REGEX=1 -DEA_HAVE_CPP11_RANDOM=1 -DEA_HAVE_CPP11_CHRONO=1 -DEA_HAVE_CPP11_SCOPED_ALLOCATOR=1 \
            -DEA_HAVE_CPP11_INITIALIZER_LIST=1 -DEA_HAVE_CPP11_SYSTEM_ERROR=1 -DEA_HAVE_CPP11_TYPEINDEX=1 -DEASTL_USER_LITERALS_ENABLED=0 \
            -DEASTL_STD_ITERATOR_CATEGORY_ENABLED=1 -DEASTL_STD_TYPE_TRAITS_AVAILABLE=1 -DEASTL_MOVE_SEMANTICS_ENABLED=1 \
            -DEASTL_VARIADIC_TEMPLATES_ENABLED=1 -DEASTL_VARIABLE_TEMPLATES_ENABLED=1 -DEASTL_INLINE_VARIABLE_ENABLED=1 \
            -DEASTL_HAVE_CPP11_TYPE_TRAITS=1 -DEASTL_INLINE_NAMESPACES_ENABLED=1 -DEASTL_ALLOCATOR_EXPLICIT_ENABLED=1 -DEA_DLL=1 \
            -DEASTL_USER_DEFINED_ALLOCATOR=1 -DEASTL_DEPRECATIONS_FOR_2024_APRIL=EA_DISABLED \
            -DLUISA_PLATFORM_UNIX -DLUISA_ENABLE_IR
# g++ 13 + the fmt bundled with spdlog: the consteval format-string check rejects LUISA_ERROR's forwarding wrapper
REF_DEFS += -DFMT_CONSTEVAL=
