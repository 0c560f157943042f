// All device code of h2j_b200 (see h2j_common.cuh for the kernel map).
#pragma once
#include "h2j_common.cuh"
#include "h2j_k_planes.cuh"
#include "h2j_k_fdct.cuh"
#include "h2j_k_fdct_fmt.cuh"
#include "h2j_k_huffman.cuh"
#include "h2j_k_entropy.cuh"
#include "h2j_k_stuff.cuh"
#include "h2j_k_pack.cuh"
