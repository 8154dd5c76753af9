// chrom.hpp — umbrella header of the C++ host layer that mirrors chrom-rs's public solver API
// (`use chrom_rs::{models::*, physics::*, solver::*}`) on top of the C ABI in chrom_cuda.h.
#pragma once
#include "chrom/models.hpp"
#include "chrom/physics.hpp"
#include "chrom/result.hpp"
#include "chrom/solver.hpp"
