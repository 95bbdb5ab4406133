// Empty stand-in for <librealsense2/rs.hpp>: /root/reference/src/types.hpp:4 includes it but the registration hot path
// (types.hpp, edge_extractor.hpp, blur_filter.hpp, *_registration.hpp, incremental_icp.hpp) uses nothing from it.
// Lets the reference's own headers be compiled on a machine that has PCL but no RealSense SDK.
#pragma once
