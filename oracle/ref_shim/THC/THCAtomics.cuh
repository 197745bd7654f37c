#pragma once
// see ../ATen/ATen.h
