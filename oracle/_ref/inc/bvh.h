#pragma once
#include "/root/reference/old/bvh copy.h"
