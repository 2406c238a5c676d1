// ORACLE SCAFFOLDING: see panmanUtils.hpp
#pragma once
#include "panmanUtils.hpp"
