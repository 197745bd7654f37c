#pragma once
// see ../ATen.h
