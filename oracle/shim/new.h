/* Memory.cpp:5 includes the pre-standard <new.h>.  TEST INFRASTRUCTURE ONLY. */
#include <new>
