#pragma once
#include "exception/all.hpp"
