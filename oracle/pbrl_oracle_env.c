#include "pbrl_oracle.h"
