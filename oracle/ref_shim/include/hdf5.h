#pragma once
typedef long hid_t; typedef int herr_t; typedef unsigned long long hsize_t; typedef int H5T_class_t;
