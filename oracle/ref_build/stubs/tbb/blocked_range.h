// ORACLE SCAFFOLDING: see _shim.h
#pragma once
#include "_shim.h"
