#pragma once
#include "/root/reference/old/raytracer_core copy.h"
