/* temporary: games not yet restated */
#include "orc.h"
const orc_game_vt orc_vt_doudizhu = {0}, orc_vt_scout = {0};
