// Empty stand-in: the reference CLIs include gtest but define no TEST()s.
#pragma once
